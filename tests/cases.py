"""Rebuild the golden cases (tests/golden/manifest.json) without the reference: seeded
weights through the drop-in module, seeded synthetic features and states, with sha256
checks against what oracle/make_golden.py recorded when it ran the reference."""
from __future__ import annotations

import gzip
import hashlib
import json
import os

import numpy as np
import torch

from deepsignal_plant_b200 import synthetic
from deepsignal_plant_b200.models import ModelBiLSTM
from oracle import model_oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MANIFEST = json.load(open(os.path.join(GOLD, "manifest.json")))
FORWARD_CASES = sorted(MANIFEST["forward"])
FEATURE_KEYS = ("kmer", "base_means", "base_stds", "base_signal_lens", "signals")


def digest(arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def build_model(entry, **extra):
    a = entry["ctor"]
    torch.manual_seed(entry["weight_seed"])
    m = ModelBiLSTM(a["seq_len"], a["signal_len"], a["num_layers1"], a["num_layers2"], a["num_classes"],
                    a["dropout_rate"], a["hidden_size"], a["vocab_size"], a["embedding_size"], a["is_base"],
                    a["is_signallen"], module=a["module"], **extra)
    m.eval()
    return m


def load_case(name, check=True, **model_extra):
    e = MANIFEST["forward"][name]
    a = e["ctor"]
    cfg = model_oracle.make_cfg(**{k: v for k, v in a.items() if k != "dropout_rate"})
    feats = synthetic.make_features(e["n"], a["seq_len"], a["signal_len"], seed=e["feature_seed"])
    states = synthetic.make_states(cfg, e["n"], seed=e["state_seed"])
    model = build_model(e, **model_extra)
    params = {k: v.detach().numpy() for k, v in model.state_dict().items()}
    if check:
        order = [g for g in ("seq", "signal", "comb") if g in states]
        assert digest(params[k] for k in params) == e["weights_sha256"], "seeded weights differ from the golden run"
        assert digest(feats[k] for k in sorted(feats)) == e["features_sha256"], "synthetic features differ"
        assert digest(x for g in order for x in states[g]) == e["states_sha256"], "synthetic states differ"
    gold = np.load(os.path.join(GOLD, "forward_%s.npz" % name))
    return dict(entry=e, cfg=cfg, feats=feats, states=states, model=model, params=params,
                logits=gold["logits"], probs=gold["probs"])


def slice_case(case, n):
    """First n sites of a case (sites are independent, models.py:178-240 has no cross-site op)."""
    out = dict(case)
    out["feats"] = {k: v[:n] for k, v in case["feats"].items()}
    out["states"] = {g: tuple(x[:, :n] for x in hc) for g, hc in case["states"].items()}
    out["logits"], out["probs"] = case["logits"][:n], case["probs"][:n]
    return out


def inject_states(model, states, device):
    """Make model.init_hidden hand out the given states in the reference's call order."""
    order = [g for g in ("seq", "signal", "comb") if g in states]

    def make():
        it = iter(order)

        def init_hidden(batch, layers, hidden):
            h0, c0 = states[next(it)]
            assert h0.shape == (layers * 2, batch, hidden)
            return (torch.from_numpy(np.ascontiguousarray(h0)).to(device),
                    torch.from_numpy(np.ascontiguousarray(c0)).to(device))
        return init_hidden
    # a fresh iterator per forward call
    class Rearm:
        def __init__(self):
            self.fn = None
            self.count = 0

        def __call__(self, batch, layers, hidden):
            if self.count % len(order) == 0:
                self.fn = make()
            self.count += 1
            return self.fn(batch, layers, hidden)
    model.init_hidden = Rearm()


def read_gz(name):
    with gzip.open(os.path.join(GOLD, name), "rt") as f:
        return f.read()
