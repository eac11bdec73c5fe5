"""bench.py contract pieces that do not need a GPU: the reference arm prints one JSON line with
the agreed keys; FLOP bookkeeping matches SURVEY.md section 8d."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "sites/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["metric"].startswith("classified sites/sec") and d["steps"] == 1 and d["scaling"] == "weak"
    # the unmodified reference from oracle/_ref when it is built (kind "reference"), else the torch-operator port
    sys.path.insert(0, ROOT)
    from oracle import ref_import
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_import.available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "sites/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_flop_bookkeeping_matches_survey():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.flops_per_site("both_bilstm", 13, 16)[0] == 118447104
    assert bench.flops_per_site("seq_bilstm", 13, 16)[0] == 126727168
    assert bench.flops_per_site("signal_bilstm", 13, 16)[0] == 127206400
    assert bench.flops_per_site("both_bilstm", 17, 20)[0] == 154950656


def test_bench_refuses_to_run_without_gpu():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                       timeout=300, cwd=ROOT, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert r.returncode != 0 and "needs a GPU" in (r.stderr + r.stdout)


def test_traffic_capture_is_tied_to_the_kernel_source():
    # roofline.traffic comes from an ncu capture; bench.py must refuse to quote it for another version of the kernels
    sys.path.insert(0, ROOT)
    import hashlib
    import bench
    t, detail = bench.traffic_of_dominant_kernel(65536)
    d = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
    sha = hashlib.sha256(open(os.path.join(ROOT, "deepsignal_plant_b200", "csrc", "kernels_tc.cu"), "rb").read()).hexdigest()
    if d.get("kernels_tc_sha256") == sha:
        assert t == d["dram_bytes_per_site"] * 65536
    else:
        assert t is None and "re-capture" in detail["reason"]


def test_command_line_object_reports_instead_of_raising():
    # bench.py's `cli` object runs tools/bench_cli.py --binary in a subprocess: with the device stubbed out it yields the
    # numbers of the host pipeline; without a GPU the real command fails and the object says so -- the bench line never
    # depends on it
    sys.path.insert(0, ROOT)
    import bench
    small = ["--block-sites", "4096"]
    d = bench.command_line_run(9000, extra=["--host-only"] + small)
    assert "unavailable" not in d and d["sites"] == 8192 == d["lines_written"] and d["value"] > 0 and d["unit"] == "sites/s"
    import torch
    if not torch.cuda.is_available():
        d = bench.command_line_run(9000, extra=small)
        assert set(d) == {"unavailable"} and "CUDA" in d["unavailable"]
