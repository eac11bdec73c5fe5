"""call_mods -> call_freq chained on the device (chain.py): the record columns computed from the classifier's
probabilities must be exactly what the reference's text round trip yields (``_call_mods`` prints, ``ModRecord``
parses), and the table aggregated from them must be byte-identical to the reference's own pipeline run through
the text file."""
import os

import numpy as np
import pytest
import torch

import cases
from deepsignal_plant_b200 import call_mods_freq as cf
from deepsignal_plant_b200 import call_modifications as cm
from deepsignal_plant_b200 import chain, synthetic
from oracle import freq_oracle, ref_import


def _probs(n, seed):
    rng = np.random.default_rng(seed)
    p1 = rng.random(n).astype(np.float32)
    p1[: n // 50] = rng.random(n // 50).astype(np.float32) * 1e-4            # scientific-notation territory ('5.6e-05')
    p1[n // 50: n // 25] = 1 - rng.random(n // 25 - n // 50).astype(np.float32) * 1e-4
    p1[-3:] = (0.0, 1.0, 0.5)
    probs = np.stack([1 - p1, p1], 1).astype(np.float32)
    return probs * (1 + rng.normal(0, 1e-7, (n, 1))).astype(np.float32)       # softmax outputs do not sum to 1 exactly


def test_records_from_probs_equals_text_round_trip_host():
    probs = _probs(100000, 0)
    a0, a1 = chain.text_round_trip(probs)
    b0, b1, lab = chain.records_from_probs(torch.from_numpy(probs))
    assert (a0 == b0.numpy()).all() and (a1 == b1.numpy()).all()
    assert (lab.numpy() == probs.argmax(1)).all()


@pytest.mark.gpu
def test_records_from_probs_equals_text_round_trip_device():
    probs = _probs(300000, 1)
    a0, a1 = chain.text_round_trip(probs)
    b0, b1, lab = chain.records_from_probs(torch.from_numpy(probs).cuda())
    assert (a0 == b0.cpu().numpy()).all() and (a1 == b1.cpu().numpy()).all()


@pytest.mark.gpu
@pytest.mark.parametrize("prob_cf,is_sort,is_bed", [(0.0, False, False), (0.004, True, False), (0.0, True, True)])
def test_device_chain_equals_the_reference_pipeline_through_text(tmp_path, prob_cf, is_sort, is_bed):
    # path A (the reference's): GPU probabilities -> the lines _call_mods prints -> file -> the UNMODIFIED reference's
    # calculate_mods_frequency + write_sitekey2stats (oracle/_ref; its restatement where that is absent);
    # path B: the same probabilities -> chain.DeviceCalls -> dsp_freq_aggregate -> the table writer.  Same bytes.
    n = 30000
    dev = torch.device("cuda:0")
    case = cases.load_case("both_13_16_s1234")
    model = cases.build_model(case["entry"], precision="fp16", max_batch=8192).cuda(0)
    feats = synthetic.make_features(n, 13, 16, seed=77)
    info = synthetic.make_sampleinfo(n, seed=77, n_chrom=4, n_pos=600)            # ~12 calls per site
    calls = chain.DeviceCalls(n, dev)
    lines = []
    fields = [s.split("\t") for s in info]
    chrom = np.array([f[0] for f in fields], dtype=object)
    pos = np.array([int(f[1]) for f in fields], np.int64)
    ids, names = cf._chrom_ids(chrom)
    keys = torch.from_numpy(cf.make_keys(ids, pos).view(np.int64)).to(dev)
    for s in range(0, n, 8192):
        e = min(n, s + 8192)
        _, probs = model(*(torch.from_numpy(feats[k][s:e]).to(dev) for k in cases.FEATURE_KEYS))
        labels = model.last_labels
        calls.append(keys[s:e], probs, labels)
        lines += cm.format_calls(info[s:e], feats["kmer"][s:e], probs.cpu().numpy(), labels.cpu().numpy())
    path = str(tmp_path / "calls.tsv")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    want_path = str(tmp_path / "want.txt")
    if ref_import.available():
        ref_freq = ref_import.import_reference("call_mods_freq")
        ref_freq.write_sitekey2stats(ref_freq.calculate_mods_frequency([path], prob_cf), want_path, is_sort, is_bed, False)
        want = open(want_path).read()
    else:
        want = freq_oracle.render(freq_oracle.aggregate(lines, prob_cf), is_sort, is_bed)
    k, p0, p1, lab = calls.columns()
    key, first, s0, s1, met, unmet, cov = (t.cpu().numpy() for t in cf._aggregate_tensors(k, p0, p1, lab, prob_cf, False, dev))
    key = key.view(np.uint64)
    rec = cf.parse_lines(lines)                                                       # text columns of the first calls only
    strand, pis, kmer = rec.meta_at(first)
    table = cf.FreqTable(np.asarray(names, dtype=object)[(key >> np.uint64(cf.POS_BITS)).astype(np.int64)],
                         (key & np.uint64((1 << cf.POS_BITS) - 1)).astype(np.int64), strand, pis, kmer, s0, s1, met, unmet, cov, first)
    got = cf.render_table(table, is_sort, is_bed)
    assert len(got) > 1000 and got == want
