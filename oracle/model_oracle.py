"""ORACLE (test infrastructure, not product code): numpy fp32 restatement of
deepsignal-plant's ``ModelBiLSTM.forward``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product package never does.

What it restates (reference paths relative to /root/reference/deepsignal_plant):

* dataflow of ``models.py:178-240`` (ModelBiLSTM.forward);
* the LSTM cell equations of the third-party ``torch.nn.LSTM`` the reference calls at
  ``models.py:196,212,226`` (torch ``nn/modules/rnn.py``: gate row order i,f,g,o;
  ``c' = sigmoid(f)*c + sigmoid(i)*tanh(g)``, ``h' = sigmoid(o)*tanh(c')``; the
  ``_reverse`` direction consumes t = T-1..0 and writes its output at its own t;
  initial states are indexed ``[layer*2 + dir]``);
* ``nn.Embedding`` (``models.py:186``), ``nn.Linear`` (``:199,215,235,238``),
  ``nn.Softmax(1)`` (``:240``).

Parity status: the reference ships no golden vectors, so this
restatement is pinned against the reference itself: ``oracle/make_golden.py`` runs the
unmodified reference module from /root/reference on CPU (fp32) with injected initial
states and stores its outputs under ``tests/golden/``; ``tests/test_oracle.py`` checks
this file against those fixtures.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def make_cfg(seq_len=13, signal_len=16, num_layers1=3, num_layers2=1, num_classes=2,
             hidden_size=256, vocab_size=16, embedding_size=4, is_base=True,
             is_signallen=True, module="both_bilstm"):
    """Shape bookkeeping of ``ModelBiLSTM.__init__`` (``models.py:103-128``)."""
    cfg = dict(seq_len=seq_len, signal_len=signal_len, num_layers1=num_layers1,
               num_layers2=num_layers2, num_classes=num_classes, hidden_size=hidden_size,
               vocab_size=vocab_size, embedding_size=embedding_size, is_base=is_base,
               is_signallen=is_signallen, module=module)
    if module == "both_bilstm":
        cfg["nhid_seq"] = hidden_size // 2
        cfg["nhid_signal"] = hidden_size - cfg["nhid_seq"]
    elif module == "seq_bilstm":
        cfg["nhid_seq"] = hidden_size
    elif module == "signal_bilstm":
        cfg["nhid_signal"] = hidden_size
    else:
        raise ValueError("--model_type is not right!")
    return cfg


def _sigmoid(x):
    return (F32(1.0) / (F32(1.0) + np.exp(-x, dtype=F32))).astype(F32)


def lstm_direction(x, w_ih, w_hh, b_ih, b_hh, h0, c0, reverse):
    """One direction of one ``nn.LSTM`` layer, batch_first.

    x (N,T,K) -> (N,T,H). Gate rows of w_ih/w_hh are [i | f | g | o] blocks of H.
    """
    N, T, _ = x.shape
    H = w_hh.shape[1]
    pre = (x.reshape(N * T, -1) @ w_ih.T).reshape(N, T, 4 * H) + (b_ih + b_hh)
    h = h0.astype(F32).copy()
    c = c0.astype(F32).copy()
    out = np.empty((N, T, H), F32)
    w_hh_t = np.ascontiguousarray(w_hh.T)
    for s in range(T):
        t = T - 1 - s if reverse else s
        g = pre[:, t, :] + h @ w_hh_t
        i = _sigmoid(g[:, 0 * H:1 * H])
        f = _sigmoid(g[:, 1 * H:2 * H])
        gg = np.tanh(g[:, 2 * H:3 * H], dtype=F32)
        o = _sigmoid(g[:, 3 * H:4 * H])
        c = f * c + i * gg
        h = o * np.tanh(c, dtype=F32)
        out[:, t, :] = h
    return out


def bilstm(x, params, prefix, num_layers, h0, c0):
    """Stacked bidirectional LSTM: layer l+1 consumes [fwd | bwd] of layer l at every t."""
    for layer in range(num_layers):
        outs = []
        for d, suffix in enumerate(("", "_reverse")):
            name = "%s.%%s_l%d%s" % (prefix, layer, suffix)
            outs.append(lstm_direction(
                x, params[name % "weight_ih"], params[name % "weight_hh"],
                params[name % "bias_ih"], params[name % "bias_hh"],
                h0[layer * 2 + d], c0[layer * 2 + d], reverse=(d == 1)))
        x = np.concatenate(outs, axis=2)
    return x


def forward(params, cfg, kmer, base_means, base_stds, base_signal_lens, signals, states):
    """ModelBiLSTM.forward -> (logits (N,C), probs (N,C)), float32.

    ``params``: state_dict as numpy float32 arrays (keys of SURVEY.md section 8a M0).
    ``states``: {'seq': (h0,c0), 'signal': (h0,c0), 'comb': (h0,c0)} -- the values the
    reference draws with ``init_hidden`` (``models.py:169-176``); explicit here because
    the reference draws fresh N(0,1) states on every call.
    """
    params = {k: np.asarray(v, F32) for k, v in params.items()}
    T = cfg["seq_len"]
    module = cfg["module"]
    N = (signals.shape[0] if module == "signal_bilstm" else np.asarray(kmer).reshape(-1, T).shape[0])
    if module != "signal_bilstm":
        cols = []
        if cfg["is_base"]:
            codes = np.asarray(kmer).reshape(N, T).astype(np.int64)
            cols.append(params["embed.weight"][codes])                    # (N,T,E)
        cols.append(np.asarray(base_means, F32).reshape(N, T, 1))
        cols.append(np.asarray(base_stds, F32).reshape(N, T, 1))
        if cfg["is_signallen"]:
            cols.append(np.asarray(base_signal_lens, F32).reshape(N, T, 1))
        x = np.concatenate(cols, axis=2).astype(F32)
        h0, c0 = states["seq"]
        out_seq = bilstm(x, params, "lstm_seq", cfg["num_layers2"], h0, c0)
        out_seq = np.maximum(out_seq @ params["fc_seq.weight"].T + params["fc_seq.bias"], F32(0))
    if module != "seq_bilstm":
        x = np.asarray(signals, F32).reshape(N, T, cfg["signal_len"])
        h0, c0 = states["signal"]
        out_sig = bilstm(x, params, "lstm_signal", cfg["num_layers2"], h0, c0)
        out_sig = np.maximum(out_sig @ params["fc_signal.weight"].T + params["fc_signal.bias"], F32(0))
    if module == "seq_bilstm":
        out = out_seq
    elif module == "signal_bilstm":
        out = out_sig
    else:
        out = np.concatenate((out_seq, out_sig), axis=2)
    h0, c0 = states["comb"]
    out = bilstm(out.astype(F32), params, "lstm_comb", cfg["num_layers1"], h0, c0)
    H = cfg["hidden_size"]
    last = np.concatenate((out[:, -1, :H], out[:, 0, H:]), axis=1)
    z = np.maximum(last @ params["fc1.weight"].T + params["fc1.bias"], F32(0))
    logits = (z @ params["fc2.weight"].T + params["fc2.bias"]).astype(F32)
    m = logits.max(axis=1, keepdims=True)
    e = np.exp(logits - m, dtype=F32)
    probs = (e / e.sum(axis=1, keepdims=True)).astype(F32)
    return logits, probs


def flops_per_site(cfg):
    """Algorithmic FLOPs (2 x MACs) of one site, shapes of SURVEY.md section 8a/8d."""
    T, H = cfg["seq_len"], cfg["hidden_size"]
    mac = 0
    module = cfg["module"]
    if module != "signal_bilstm":
        kin = (cfg["embedding_size"] if cfg["is_base"] else 0) + (3 if cfg["is_signallen"] else 2)
        hs = cfg["nhid_seq"]
        mac += 2 * T * 4 * hs * (kin + hs) + T * 2 * hs * hs
    if module != "seq_bilstm":
        hs = cfg["nhid_signal"]
        mac += 2 * T * 4 * hs * (cfg["signal_len"] + hs) + T * 2 * hs * hs
    for layer in range(cfg["num_layers1"]):
        kin = H if layer == 0 else 2 * H
        mac += 2 * T * 4 * H * (kin + H)
    mac += 2 * H * H + H * cfg["num_classes"]
    return 2 * mac
