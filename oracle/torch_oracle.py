"""ORACLE (test infrastructure, not product code): the reference forward restated on torch's own
CPU operators.

``deepsignal_plant/models.py:178-240`` calls ``nn.Embedding``, ``nn.LSTM`` (oneDNN
``mkldnn_rnn_layer`` on CPU), ``nn.Linear`` and ``nn.Softmax``; this module drives the same torch
CPU kernels through ``torch.nn.functional`` / ``torch._VF.lstm`` from a plain ``state_dict``, so
that the CPU baseline timed next to the GPU numbers (``bench.py`` ``cpu_baseline`` /
``--impl reference``) is the reference's actual CPU code path -- torch, all host threads -- and not
a numpy stand-in.  ``call_mods_batches`` adds the batch loop and random initial states of
``_call_mods`` (``call_modifications.py:147-169``, ``models.py:169-176``).  Pinned by the same
fixtures as ``model_oracle`` (``tests/test_oracle.py``)."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _lstm(x, sd, prefix, layers, h0, c0):
    """``nn.LSTM(batch_first=True, bidirectional=True)`` forward from state_dict entries."""
    flat = []
    for l in range(layers):
        for sfx in ("", "_reverse"):
            flat += [sd["%s.weight_ih_l%d%s" % (prefix, l, sfx)], sd["%s.weight_hh_l%d%s" % (prefix, l, sfx)],
                     sd["%s.bias_ih_l%d%s" % (prefix, l, sfx)], sd["%s.bias_hh_l%d%s" % (prefix, l, sfx)]]
    out, _, _ = torch._VF.lstm(x, (h0, c0), flat, True, layers, 0.0, False, True, True)
    return out


def forward(sd, cfg, kmer, base_means, base_stds, base_signal_lens, signals, states):
    """sd: state_dict of CPU float32 tensors; cfg: ``model_oracle.make_cfg``; inputs: CPU tensors;
    states: {"seq"|"signal"|"comb": (h0, c0)} tensors shaped (layers*2, N, hidden).  -> (logits, probs)."""
    T, H, mod = cfg["seq_len"], cfg["hidden_size"], cfg["module"]
    with torch.no_grad():
        parts = []
        if mod != "signal_bilstm":
            cols = []
            if cfg["is_base"]:
                cols.append(F.embedding(kmer.long(), sd["embed.weight"]))                      # models.py:186
            cols += [base_means.reshape(-1, T, 1).float(), base_stds.reshape(-1, T, 1).float()]
            if cfg["is_signallen"]:
                cols.append(base_signal_lens.reshape(-1, T, 1).float())
            x = torch.cat(cols, 2)                                                             # :188-195
            x = _lstm(x, sd, "lstm_seq", cfg["num_layers2"], *states["seq"])                  # :196
            parts.append(F.relu(F.linear(x, sd["fc_seq.weight"], sd["fc_seq.bias"])))          # :199-201
        if mod != "seq_bilstm":
            x = _lstm(signals.float(), sd, "lstm_signal", cfg["num_layers2"], *states["signal"])   # :212
            parts.append(F.relu(F.linear(x, sd["fc_signal.weight"], sd["fc_signal.bias"])))    # :215-217
        x = parts[0] if len(parts) == 1 else torch.cat(parts, 2)                               # :220-225
        x = _lstm(x, sd, "lstm_comb", cfg["num_layers1"], *states["comb"])                    # :226
        x = torch.cat((x[:, -1, :H], x[:, 0, H:]), 1)                                          # :229-231
        x = F.relu(F.linear(x, sd["fc1.weight"], sd["fc1.bias"]))                              # :234-237
        logits = F.linear(x, sd["fc2.weight"], sd["fc2.bias"])                                 # :238
        return logits, F.softmax(logits, 1)                                                    # :240


def call_mods_batches(sd, cfg, feats, batch_size=512):
    """The reference's inference loop over one feature block: slices of ``batch_size``, fresh
    ``torch.randn`` states per forward (``models.py:169-176``), argmax labels
    (``call_modifications.py:147-169``).  feats: dict of CPU tensors.  Returns sites processed."""
    n = feats["signals"].shape[0] if "signals" in feats else feats["kmer"].shape[0]
    groups = (("seq", cfg["num_layers2"], cfg.get("nhid_seq", 0)), ("signal", cfg["num_layers2"], cfg.get("nhid_signal", 0)),
              ("comb", cfg["num_layers1"], cfg["hidden_size"]))
    for s in range(0, n, batch_size):
        e = min(s + batch_size, n)
        states = {g: (torch.randn(l * 2, e - s, h), torch.randn(l * 2, e - s, h)) for g, l, h in groups if h}
        _, probs = forward(sd, cfg, feats["kmer"][s:e], feats["base_means"][s:e], feats["base_stds"][s:e],
                           feats["base_signal_lens"][s:e], feats["signals"][s:e], states)
        torch.max(probs, 1)
    return n
