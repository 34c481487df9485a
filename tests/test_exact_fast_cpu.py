"""CPU check of the parallel exact member sums (quant_b200/csrc/qb200_exact_fast.cuh): the per-thread bodies the CUDA
kernels run are compiled for the host and compared, bit for bit, with the reference's compensated loop
(Solution::sumInArea, /root/reference/src/Quantizer.cpp:59-70) in IEEE double - tests/cpp/exact_fast_test.cpp."""
import os
import subprocess

from conftest import ROOT


def test_parallel_decomposition_reproduces_the_compensated_loop_bit_for_bit(tmp_path):
    exe = str(tmp_path / "exact_fast_test")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-I",
                           os.path.join(ROOT, "quant_b200", "csrc"), os.path.join(ROOT, "tests", "cpp", "exact_fast_test.cpp"),
                           "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    assert " 0 mismatches" in r.stdout.splitlines()[-1]
