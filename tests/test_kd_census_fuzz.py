"""Short run of the CPU fuzz of the auto mode's census (tools/fuzz_kd_census.py): the KD tree built from integer-sum
centroids must have the shape of the tree built from the reference's compensated-sum centroids whenever the census
reports a robust margin, and ties it calls order-safe must be decided as the reference's walk decides them."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_census_fuzz_short():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_kd_census.py"), "8", "5"], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 of them with a different shape" in r.stdout and "0 of them decided differently" in r.stdout
