"""tools/fast5_to_archive.py against a stand-in for h5py (h5py itself is not in this image): the converter
must read the fields the reference's accessors read (extract_features.py:44-176, 255-273) and write an
archive that round-trips.  Where /root/reference is present (the build container) the reference's own
accessors are run on the same stand-in and must agree."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest

from deepsignal_plant_b200 import extract_features as ef, synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


class _Node(dict):
    """A group or dataset: dict of children, .attrs, [()] / field access for datasets."""

    def __init__(self, children=None, attrs=None, data=None):
        super().__init__(children or {})
        self.attrs, self._data = attrs or {}, data

    def __getitem__(self, key):
        if self._data is not None:
            return self._data if isinstance(key, tuple) and key == () else self._data[key]
        node = self
        for part in [p for p in key.split("/") if p]:
            node = dict.__getitem__(node, part)
        return node

    def __contains__(self, key):
        try:
            self[key]
            return True
        except KeyError:
            return False

    def astype(self, t):
        return self._data.astype(t)


class _File(_Node):
    store = {}

    def __init__(self, path, mode="r"):
        if path not in self.store:
            raise IOError("cannot open " + path)
        root = self.store[path]
        super().__init__(dict(root), root.attrs)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def close(self):
        pass


def fake_fast5(rd, with_alignment=True, with_channel=True):
    rel = 37
    ev = np.zeros(len(rd["ev_len"]), dtype=[("start", "<i8"), ("length", "<i8"), ("base", "S1")])
    ev["start"], ev["length"] = np.asarray(rd["ev_start"]) - rel, rd["ev_len"]
    ev["base"] = [c.encode() for c in rd["ev_base"]]
    events = _Node(attrs={"read_start_rel_to_raw": np.int64(rel)}, data=ev)
    tmpl = {"Events": events}
    if with_alignment:
        tmpl["Alignment"] = _Node(attrs={"mapped_strand": rd["alignstrand"].encode(), "mapped_chrom": rd["chrom"].encode(),
                                         "mapped_start": np.int64(rd["chrom_start"])})
    root = {"Analyses": _Node({"RawGenomeCorrected_000": _Node({"BaseCalled_template": _Node(tmpl)})}),
            "Raw": _Node({"Reads": _Node({"Read_17": _Node({"Signal": _Node(data=rd["raw"])},
                                                           attrs={"read_id": rd["readname"].encode()})})})}
    if with_channel:
        digi, rng = np.float64(8192.0), np.float64(rd["scaling"]) * 8192.0
        root["UniqueGlobalKey"] = _Node({"channel_id": _Node(attrs={"digitisation": digi, "range": rng, "offset": np.float64(rd["offset"])})})
    return _Node(root)


@pytest.fixture
def converter(monkeypatch):
    fake = types.ModuleType("h5py")
    fake.File = _File
    monkeypatch.setitem(sys.modules, "h5py", fake)
    spec = importlib.util.spec_from_file_location("fast5_to_archive", os.path.join(ROOT, "tools", "fast5_to_archive.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _File.store = {}
    return mod


def test_converter_reads_what_the_reference_accessors_read(converter, tmp_path, monkeypatch):
    reads = synthetic.make_reads(5, seed=3, mean_bases=80)
    d = tmp_path / "f5" / "sub"
    d.mkdir(parents=True)
    for i, rd in enumerate(reads):
        p = str(d / ("read%d.fast5" % i))
        open(p, "w").close()
        _File.store[p] = fake_fast5(rd, with_alignment=i != 3, with_channel=i != 1)
    bad = str(d / "broken.fast5")
    open(bad, "w").close()                                  # not in the store: opening it raises, the file is skipped
    dec = lambda i: converter.decode_fast5(str(d / ("read%d.fast5" % i)), "RawGenomeCorrected_000", "BaseCalled_template")
    got = {i: dec(i) for i in (0, 2, 3, 4)}
    assert got[3] is None                                   # no Alignment group: the reference returns empty strings and moves on
    with pytest.raises(KeyError):                           # no channel info: an error in the reference too (see below)
        dec(1)
    for i in (0, 2, 4):
        rd, g = reads[i], got[i]
        for k in ("readname", "strand", "alignstrand", "chrom", "chrom_start", "ev_base"):
            assert g[k] == rd[k], k
        assert np.array_equal(g["raw"], rd["raw"]) and g["raw"].dtype == np.int16
        assert np.array_equal(g["ev_start"], rd["ev_start"]) and np.array_equal(g["ev_len"], rd["ev_len"])
        assert g["scaling"] == np.float64(rd["scaling"]) * 8192.0 / np.float64(8192.0) and g["offset"] == rd["offset"]
    out = str(tmp_path / "reads.npz")
    monkeypatch.setattr(sys, "argv", ["fast5_to_archive.py", "-i", str(tmp_path / "f5"), "-o", out])
    converter.main()
    batch = ef.load_reads(out)
    assert sorted(batch.readname) == sorted(reads[i]["readname"] for i in (0, 2, 4))
    if os.path.isdir(REF):                                  # build container: the reference's accessors on the same stand-in
        sys.path.insert(0, os.path.join(ROOT))
        from oracle.make_golden import import_reference
        import_reference()
        import deepsignal_plant.extract_features as ref_ex
        monkeypatch.setattr(ref_ex, "h5py", sys.modules["h5py"])
        for i in (0, 2, 4):
            p = str(d / ("read%d.fast5" % i))
            g = got[i]
            assert ref_ex._get_alignment_info_from_fast5(p, "RawGenomeCorrected_000", "BaseCalled_template") == (
                g["readname"], g["strand"], g["alignstrand"], g["chrom"], g["chrom_start"])
            raw, events = ref_ex._get_label_raw(p, "RawGenomeCorrected_000", "BaseCalled_template")
            assert np.array_equal(raw, g["raw"])
            assert [int(e[0]) for e in events] == g["ev_start"].tolist() and [int(e[1]) for e in events] == g["ev_len"].tolist()
            assert "".join(e[2] for e in events) == g["ev_base"]
            assert ref_ex._get_scaling_of_a_read(p) == (g["scaling"], g["offset"])
        with pytest.raises(KeyError):
            ref_ex._get_scaling_of_a_read(str(d / "read1.fast5"))
        assert ref_ex._get_alignment_info_from_fast5(str(d / "read3.fast5"), "RawGenomeCorrected_000", "BaseCalled_template") == ("",) * 5
