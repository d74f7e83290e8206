"""CPU: the resident helper of the drop-in executables (necat_b200/host/nb_helper.h, opt-in with NECAT_B200_HELPER=1) on the
emulated build (tests/emu): same files and exit status as the stand-alone run, an error inside a job reaches the caller,
the helper leaves after its idle time and removes its socket."""
import os
import subprocess
import time

import pytest

import util

EMU_DIR = os.path.join(util.ROOT, "tests", "emu")
EMU_BIN = os.path.join(EMU_DIR, "_build", "cudaemu", "bin")
PM11 = "-k 11 -z 10 -q 100 -b 2000 -s 3 -n 500 -a 1000 -d 0.25 -e 0.5 -m 500 -t 1 -j 0 -u 1 -i 1"
CNS = "-a 1000 -x 3 -y 12 -l 500 -e 0.5 -p 0.8 -u 0 -r 0 -f 0"


@pytest.fixture(scope="module")
def emu(built):
    subprocess.check_call(["make", "-s", "-C", EMU_DIR])


def _flow(tmp_path, wrk, tag, env):
    e = dict(os.environ, **env)
    cand = str(tmp_path / f"{tag}.bin")
    p = subprocess.run([os.path.join(EMU_BIN, "oc2pmov")] + PM11.split() + [wrk, "0", cand], env=e, capture_output=True, text=True, cwd=str(tmp_path))
    assert p.returncode == 0, p.stderr
    subprocess.run([os.path.join(EMU_BIN, "oc2pcan"), "-t", "1", wrk, cand], env=e, check=True, capture_output=True)
    p = subprocess.run([os.path.join(EMU_BIN, "oc2cns")] + CNS.split() + [wrk, cand, f"{tag}_cns.fa", f"{tag}_raw.fa"], env=e,
                       capture_output=True, text=True, cwd=str(tmp_path))          # relative output paths: the job runs in the caller's directory
    assert p.returncode == 0, p.stderr
    return open(cand, "rb").read(), open(tmp_path / f"{tag}_cns.fa").read(), open(tmp_path / f"{tag}_raw.fa").read(), p.stdout + p.stderr


def test_helper_runs_the_same_jobs(emu, tmp_path):
    wrk, _ = util.make_workdir(tmp_path, genome=6000, depth=10, seed=93, err=0.10)
    plain = _flow(tmp_path, wrk, "plain", {})
    henv = {"NECAT_B200_HELPER": "1", "NECAT_B200_HELPER_IDLE": "2", "NECAT_B200_HELPER_DIR": str(tmp_path)}
    first = _flow(tmp_path, wrk, "h1", henv)         # starts the helper
    second = _flow(tmp_path, wrk, "h2", henv)        # finds it
    assert plain[:3] == first[:3] == second[:3] and len(plain[1]) > 1000
    assert "consensus partition 0" in second[3]                          # the banners come back through the caller's descriptors
    socks = [f for f in os.listdir(tmp_path) if f.endswith(".sock")]
    assert len(socks) == 1
    # an error inside a job: exit status 1 and the message on the caller's stderr; the next call starts a new helper
    p = subprocess.run([os.path.join(EMU_BIN, "oc2pmov")] + PM11.split() + [wrk, "7", str(tmp_path / "x.bin")], env=dict(os.environ, **henv),
                       capture_output=True, text=True)
    assert p.returncode == 1 and "out of range" in p.stderr
    assert _flow(tmp_path, wrk, "h3", henv)[:3] == plain[:3]
    # other tunables, other helper (the library reads NB_* once per process)
    assert _flow(tmp_path, wrk, "h4", dict(henv, NB_CNS_SPAN="64"))[:3] == plain[:3]
    assert [f for f in os.listdir(tmp_path) if f.endswith(".sock") and f not in socks], "a job with other tunables went to the same helper"
    time.sleep(3.5)
    assert not [f for f in os.listdir(tmp_path) if f.endswith(".sock")], "the helpers did not leave after their idle time"


def test_helper_is_opt_in(emu, tmp_path):
    wrk, _ = util.make_workdir(tmp_path, genome=6000, depth=10, seed=94, err=0.10)
    _flow(tmp_path, wrk, "plain", {"NECAT_B200_HELPER_DIR": str(tmp_path)})
    assert not [f for f in os.listdir(tmp_path) if f.endswith(".sock")]
