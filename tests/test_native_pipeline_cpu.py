"""CPU: the C++ host layer end to end -- database in, pair lists, uploads, gpu_index slicing, write order and
write rules, resume, sequential ordering by name, verify_matches -- against a CPU stand-in for libb200match.so
(tests/helpers/mock_b200match.cpp: oracle matcher + placeholder verifier), put in front of the real library
with LD_LIBRARY_PATH inside a subprocess.  Test infrastructure only; the product never sees the mock."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def mock_dir():
    out = os.path.join(HERE, "helpers", "mock")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libb200match.so")
    src = os.path.join(HERE, "helpers", "mock_b200match.cpp")
    hdr = os.path.join(ROOT, "include", "b200match.h")
    oracle_dir = os.path.join(ROOT, "oracle")
    subprocess.check_call(["make", "-C", oracle_dir, "-s"])
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-Wall", "-o", so, src,
                               "-L", oracle_dir, "-loracle", f"-Wl,-rpath,{oracle_dir}"])
    return out


def test_cxx_pipelines_against_the_mock_library(mock_dir, tmp_path):
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = mock_dir + os.pathsep + env.get("LD_LIBRARY_PATH", "")
    r = subprocess.run([sys.executable, os.path.join(HERE, "helpers", "native_pipeline_script.py"), str(tmp_path)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "NATIVE-PIPELINE-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
