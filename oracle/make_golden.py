"""Generate tests/golden/* by RUNNING THE REFERENCE (test infrastructure, not product code).

Run here (the build container), never on the GPU box: it imports the unmodified reference
from /root/reference (read-only), which does not travel.

    python oracle/make_golden.py            # writes tests/golden/

What it pins (SURVEY.md section 8c -- the reference ships no golden vectors of its own):

1. forward_<case>.npz -- ``deepsignal_plant.models.ModelBiLSTM`` (models.py:99-240) on CPU
   fp32 with seeded random-init weights, seeded synthetic features and explicit initial
   states injected by overriding ``init_hidden`` on the reference instance; stores logits,
   probs and sha256 digests of weights / features / states so a test can prove it rebuilt
   the same inputs.
2. callmods_<seed>.* -- output lines of the reference's ``_call_mods``
   (call_modifications.py:130-192) plus the raw float32 probabilities it formatted.
3. freq_<case>.txt.gz -- bytes written by the reference's ``calculate_mods_frequency`` +
   ``write_sitekey2stats`` (call_mods_freq.py:29-122) for synthetic call_mods inputs.

It also checks the numpy/Python restatements under oracle/ against the reference while it
has both in hand, and refuses to write fixtures if they disagree.
"""
from __future__ import annotations

import gzip
import hashlib
import json
import os
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
GOLD = os.path.join(ROOT, "tests", "golden")

from deepsignal_plant_b200 import synthetic  # noqa: E402
from oracle import model_oracle, callmods_oracle, freq_oracle  # noqa: E402

# (name, ctor kwargs, weight seed, feature seed, n sites)
FORWARD_CASES = [
    ("both_13_16_s1234", dict(), 1234, 0, 10000),
    ("both_13_16_s1", dict(), 1, 11, 10000),
    ("both_13_16_s2", dict(), 2, 12, 10000),
    ("seq_13_16_s1234", dict(module="seq_bilstm"), 1234, 3, 4096),
    ("signal_13_16_s1234", dict(module="signal_bilstm"), 1234, 4, 4096),
    ("both_17_20_s1234", dict(seq_len=17, signal_len=20), 1234, 5, 4096),
    # small odd shapes: ragged batch, 2-layer branches, no base / no signal-length features
    ("both_small_odd", dict(seq_len=5, signal_len=8, num_layers1=2, num_layers2=2, hidden_size=64), 7, 6, 257),
    ("seq_nobase", dict(seq_len=9, hidden_size=32, num_layers1=1, is_base=False, module="seq_bilstm"), 8, 7, 77),
    ("both_nolen_c3", dict(seq_len=7, signal_len=12, hidden_size=96, num_classes=3, is_signallen=False), 9, 8, 130),
]
STATE_SEED = 4321


def import_reference():
    """Import the reference package; h5py / statsmodels are absent here and only used by
    the fast5 extraction path (extract_features.py:13,24), so stub them."""
    for name in ("h5py", "statsmodels"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["statsmodels"].robust = types.ModuleType("statsmodels.robust")
    sys.modules["statsmodels.robust"] = sys.modules["statsmodels"].robust
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import deepsignal_plant.models as ref_models
    import deepsignal_plant.call_modifications as ref_cm
    import deepsignal_plant.call_mods_freq as ref_freq
    return ref_models, ref_cm, ref_freq


def digest(arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def ctor_args(kw):
    d = dict(seq_len=13, signal_len=16, num_layers1=3, num_layers2=1, num_classes=2, dropout_rate=0,
             hidden_size=256, vocab_size=16, embedding_size=4, is_base=True, is_signallen=True,
             module="both_bilstm")
    d.update(kw)
    return d


def build_inputs(kw, wseed, fseed, n):
    """Everything a test needs to rebuild the case without the reference."""
    a = ctor_args(kw)
    cfg = model_oracle.make_cfg(**{k: v for k, v in a.items() if k != "dropout_rate"})
    feats = synthetic.make_features(n, a["seq_len"], a["signal_len"], seed=fseed)
    states = synthetic.make_states(cfg, n, seed=STATE_SEED)
    return a, cfg, feats, states


def forward_goldens(ref_models):
    manifest = {}
    for name, kw, wseed, fseed, n in FORWARD_CASES:
        a, cfg, feats, states = build_inputs(kw, wseed, fseed, n)
        torch.manual_seed(wseed)
        model = ref_models.ModelBiLSTM(a["seq_len"], a["signal_len"], a["num_layers1"], a["num_layers2"],
                                       a["num_classes"], a["dropout_rate"], a["hidden_size"], a["vocab_size"],
                                       a["embedding_size"], a["is_base"], a["is_signallen"], module=a["module"])
        model.eval()
        order = [g for g in ("seq", "signal", "comb") if g in states]
        calls = iter(order)

        def injected(batch, layers, hidden, _calls=calls):
            h0, c0 = states[next(_calls)]
            assert h0.shape == (layers * 2, batch, hidden)
            return torch.from_numpy(h0), torch.from_numpy(c0)
        model.init_hidden = injected
        with torch.no_grad():
            logits, probs = model(*(torch.from_numpy(feats[k]) for k in
                                    ("kmer", "base_means", "base_stds", "base_signal_lens", "signals")))
        logits, probs = logits.numpy(), probs.numpy()
        params = {k: v.numpy() for k, v in model.state_dict().items()}
        o_logits, o_probs = model_oracle.forward(params, cfg, feats["kmer"], feats["base_means"], feats["base_stds"],
                                                 feats["base_signal_lens"], feats["signals"], states)
        err = float(np.abs(o_probs - probs).max())
        agree = float((o_probs.argmax(1) == probs.argmax(1)).mean())
        print("%-22s n=%5d  oracle-vs-reference max|dprob|=%.2e labels=%.4f%%  prob1 mean=%.4f sd=%.4f"
              % (name, n, err, agree * 100, probs[:, 1].mean(), probs[:, 1].std()))
        assert err < 5e-6, "numpy oracle disagrees with the reference"
        entry = dict(ctor=a, weight_seed=wseed, feature_seed=fseed, state_seed=STATE_SEED, n=n,
                     weights_sha256=digest(params[k] for k in params),
                     features_sha256=digest(feats[k] for k in sorted(feats)),
                     states_sha256=digest(x for g in order for x in states[g]),
                     oracle_max_abs_dprob=err, oracle_label_agreement=agree)
        manifest[name] = entry
        np.savez_compressed(os.path.join(GOLD, "forward_%s.npz" % name), logits=logits, probs=probs)
    return manifest


def callmods_golden(ref_models, ref_cm, seed=77, n=3000, batch=512):
    a, cfg, feats, _ = build_inputs({}, 1234, 21, n)
    torch.manual_seed(1234)
    model = ref_models.ModelBiLSTM(13, 16, 3, 1, 2, 0, 256, 16, 4, True, True, module="both_bilstm")
    model.eval()
    info = synthetic.make_sampleinfo(n, seed=21)
    kmers = feats["kmer"].astype(np.int64).tolist()
    batch_lists = (info, kmers, feats["base_means"].tolist(), feats["base_stds"].tolist(),
                   feats["base_signal_lens"].astype(np.int64).tolist(), feats["signals"].tolist(),
                   [0] * n)
    seen = []

    class Recorder:
        def __call__(self, *args):
            out = model(*args)
            seen.append(out[1].detach().numpy().copy())
            return out
    torch.manual_seed(seed)          # the reference draws h0/c0 from the global CPU generator
    lines, acc, nb = ref_cm._call_mods(batch_lists, Recorder(), batch)
    probs = np.concatenate(seen, 0)
    o_lines, o_labels = callmods_oracle.call_lines(info, feats["kmer"], probs)
    assert o_lines == lines, "callmods oracle disagrees with the reference"
    with gzip.open(os.path.join(GOLD, "callmods_%d.tsv.gz" % seed), "wt") as f:
        f.write("\n".join(lines) + "\n")
    np.savez_compressed(os.path.join(GOLD, "callmods_%d_probs.npz" % seed), probs=probs)
    print("callmods golden: %d lines, %d batches, accuracy-vs-zero-labels %.4f" % (len(lines), nb, acc))
    return dict(n=n, batch=batch, weight_seed=1234, feature_seed=21, rng_seed=seed,
                lines_sha256=hashlib.sha256(("\n".join(lines)).encode()).hexdigest())


EDGE_LINES = """\
chr2\t100\t+\t100\tr1\tt\t0.9\t0.1\t0\tAACGT
chr10\t5\t-\t994\tr1\tt\t0.2\t0.8\t1\tTTCGA
chr2\t100\t-\t899\tr2\tt\t0.1\t0.9\t1\tGGCAT
chr2\t7\t+\t7\tr2\tt\t0.5\t0.5\t0\tAACAA
chr2\t7\t-\t992\tr3\tt\t0.45\t0.55\t1\tCCCGG
chr2\t7\t+\t7\tr4\tt\t0.7\t0.3\t0\tAACAA
chr1\t1\t+\t1\tr4\tt\t5.6e-05\t0.999944\t1\tACCGT
chr1\t1\t+\t1\tr5\tt\t1e-06\t0.999999\t1\tACCGT
chr1\t1\t+\t1\tr6\tt\t0.0\t1.0\t1\tACCGT
chr1\t1\t+\t1\tr7\tt\t1.0\t0.0\t0\tACCGT
chr1\t2\t+\t2\tr7\tt\t0.0005\t0.9995\t1\tGGCTT
chr1\t2\t+\t2\tr8\tt\t0.001\t0.999\t1\tGGCTT
chr1\t2\t+\t2\tr9\tt\t0.001\t0.999\t1\tGGCTT
chr1\t3\t+\t3\tr9\tt\t0.3335\t0.6665\t1\tTTCAA
chr1\t3\t+\t3\tr10\tt\t0.3335\t0.6665\t1\tTTCAA
chr1\t3\t+\t3\tr11\tt\t0.3335\t0.6665\t0\tTTCAA
chrX\t12\t-\t30\tr11\tt\t0.125\t0.875\t1\tAGCTA
chrX\t12\t-\t30\tr12\tt\t0.0625\t0.9375\t1\tAGCTA
"""


def freq_goldens(ref_freq):
    manifest = {}
    big = synthetic.make_callmods_records(100000, n_chrom=12, n_pos=900, seed=5)
    inputs = {"edge": EDGE_LINES.splitlines(), "synth": big}
    with open(os.path.join(GOLD, "freq_edge_input.tsv"), "w") as f:
        f.write(EDGE_LINES)
    cases = [("edge", 0.0, False, False), ("edge", 0.0, True, False), ("edge", 0.1, False, False),
             ("edge", 0.0, True, True), ("edge", 0.5, False, True),
             ("synth", 0.0, False, False), ("synth", 0.0, True, False), ("synth", 0.5, False, False),
             ("synth", 0.0, True, True), ("synth", 0.5, False, True)]
    with tempfile.TemporaryDirectory() as tmp:
        paths = {}
        for key, lines in inputs.items():
            # split across two files: the reference concatenates files in argument order
            half = len(lines) // 2
            p1, p2 = os.path.join(tmp, key + "_a.tsv"), os.path.join(tmp, key + "_b.tsv.gz")
            with open(p1, "w") as f:
                f.write("\n".join(lines[:half]) + "\n")
            with gzip.open(p2, "wt") as f:
                f.write("\n".join(lines[half:]) + "\n")
            paths[key] = [p1, p2]
            manifest["input_" + key] = dict(n=len(lines), sha256=hashlib.sha256("\n".join(lines).encode()).hexdigest())
        for key, cf, is_sort, is_bed in cases:
            stats = ref_freq.calculate_mods_frequency(paths[key], cf)
            out = os.path.join(tmp, "out.txt")
            ref_freq.write_sitekey2stats(stats, out, is_sort, is_bed, False)
            data = open(out).read()
            table = freq_oracle.aggregate(inputs[key], cf)
            assert freq_oracle.render(table, is_sort, is_bed) == data, "freq oracle disagrees with the reference"
            name = "freq_%s_cf%s_%s_%s" % (key, str(cf).replace(".", "p"), "sorted" if is_sort else "unsorted",
                                           "bed" if is_bed else "tsv")
            with gzip.open(os.path.join(GOLD, name + ".txt.gz"), "wt") as f:
                f.write(data)
            manifest[name] = dict(input=key, prob_cf=cf, sort=is_sort, bed=is_bed, n_sites=data.count("\n"))
            print("%-40s %6d sites" % (name, data.count("\n")))
    return manifest


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    ref_models, ref_cm, ref_freq = import_reference()
    manifest = {"torch": torch.__version__, "numpy": np.__version__,
                "note": "generated by oracle/make_golden.py from the unmodified reference at /root/reference"}
    manifest["forward"] = forward_goldens(ref_models)
    manifest["callmods"] = callmods_golden(ref_models, ref_cm)
    manifest["freq"] = freq_goldens(ref_freq)
    with open(os.path.join(GOLD, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print("wrote", GOLD)


if __name__ == "__main__":
    main()
