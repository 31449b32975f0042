"""CPU tests of the host logic: the reference's own host pass (unmodified x3.c,
dict.c, ... compiled into oracle/_ref by oracle/Makefile) driven by the NEW backend
shim must emit the reference's stream.  On a CPU-only machine the device layer is
the oracle-backed fake (oracle/ref_hooks/fake_search_oracle.c, test-only); the GPU
version of this test lives in test_gpu_parity.py."""
import hashlib
import json
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / "oracle" / "_ref"
KATS = json.loads((ROOT / "tests" / "golden" / "streams.json").read_text())

CASES = [k for k in KATS if k.split(":")[0] in ("C4", "C5") or k.split(":")[2] in ("", "-n 3 -t 7", "-t 3 -w 1")]


@pytest.mark.skipif(not (REF / "x3_ref_dropin_cpuoracle").exists(), reason="oracle/_ref not built")
@pytest.mark.parametrize("case", sorted(CASES))
def test_reference_host_over_new_shim_matches_kat(case, corpus, tmp_path):
    name, size, flags = case.split(":")
    data = corpus.generate(name, int(size))
    assert hashlib.sha256(data).hexdigest() == KATS[case]["in_sha256"], "corpus generator drifted"
    src = tmp_path / "in.bin"
    src.write_bytes(data)
    out = tmp_path / "out.x3"
    subprocess.run([str(REF / "x3_ref_dropin_cpuoracle"), "-zf", *flags.split(), str(src), str(out)], check=True,
                   stderr=subprocess.DEVNULL)
    s = out.read_bytes()
    assert len(s) == KATS[case]["len"]
    assert hashlib.sha256(s).hexdigest() == KATS[case]["sha256"]


@pytest.mark.skipif(not (REF / "x3_ref_check_cpuoracle").exists(), reason="oracle/_ref not built")
@pytest.mark.parametrize("flags", ["", "-m 1 -n 1", "-t 50 -w 32"])
def test_every_call_equals_reference(flags, corpus, tmp_path):
    """Interposition harness: reference and new find_best_match side by side on every call."""
    src = tmp_path / "in.bin"
    src.write_bytes(corpus.generate("C1", 30000))
    r = subprocess.run([str(REF / "x3_ref_check_cpuoracle"), "-zf", *flags.split(), str(src), str(tmp_path / "o.x3")],
                       stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr[-400:]
    assert "mismatches 0" in r.stderr
    # and the reference decoder restores the input
    subprocess.run([str(REF / "x3_ref"), "-df", str(tmp_path / "o.x3"), str(tmp_path / "back")], check=True,
                   stderr=subprocess.DEVNULL)
    assert (tmp_path / "back").read_bytes() == src.read_bytes()
