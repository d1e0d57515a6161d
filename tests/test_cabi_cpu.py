"""CPU-side checks of the product's host logic and C ABI (no GPU compute):
the shared library loads and exports every symbol include/formoniq_b200.h
declares, the tape compiler reproduces the oracle bit for bit, the closed-form
Kuhn numbering equals the reference's sort+dedup numbering, and compute entry
points fail loudly without a device."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def fq():
    from formoniq_b200 import build

    build.build()
    import formoniq_b200

    return formoniq_b200


def test_library_exports_every_declared_symbol(fq):
    from formoniq_b200 import _lib

    header = open(os.path.join(ROOT, "include", "formoniq_b200.h")).read()
    declared = set(re.findall(r"\b(fq_[a-z0-9_]+)\s*\(", header))
    declared -= {"fq_ctx", "fq_mesh", "fq_csr", "fq_vec", "fq_kind"}
    bound = {name for name, _, _ in _lib.SIGNATURES}
    assert declared == bound, (declared - bound, bound - declared)
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), name


def test_tape_compiler_matches_oracle_bitwise(tmp_path):
    exe = tmp_path / "tape_host_check"
    subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-ffp-contract=off",
                           os.path.join(ROOT, "tests", "cpp", "tape_host_check.cpp"), "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert out.stdout.startswith("OK "), out.stdout
    assert int(out.stdout.split()[1]) > 10000


def test_tile_plan_host_builder_and_staged_sets_reproduce_the_oracle(tmp_path):
    # the tile-fused kernel's plan (row slots, record streams) built by the host reference builder and interpreted on
    # the CPU with the GENERATED staged block-set functions must give the oracle's structural pattern, its values
    # bitwise, and its value-dependent (`!= 0.0`) pattern, for every Hodge block set and single block, dims 1..3
    from formoniq_b200 import build as B

    B.generate()
    exe = tmp_path / "tile_plan_check"
    subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-ffp-contract=off",
                           os.path.join(ROOT, "tests", "cpp", "tile_plan_check.cpp"), "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert out.stdout.startswith("OK "), out.stdout
    ncases, nnz = map(int, out.stdout.split()[1:3])
    assert ncases >= 100 and nnz > 100000


def test_host_widening_of_downloaded_indices(tmp_path):
    # u32 -> usize in place on host threads (csrc/host_widen.hpp): what fq_csr_download_async does to the index arrays
    exe = tmp_path / "host_widen_check"
    subprocess.check_call(["/usr/bin/g++", "-O3", "-std=c++17", "-pthread",
                           os.path.join(ROOT, "tests", "cpp", "host_widen_check.cpp"), "-o", str(exe)])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert out.stdout.startswith("OK "), out.stdout


def test_kuhn_closed_form_numbering_matches_reference_construction(fq):
    from oracle import oracle as O

    for dim, shape in ((1, [3]), (2, [2, 3]), (2, [4, 4]), (3, [2, 3, 2]), (3, [3, 3, 3]), (4, [2, 1, 2, 2]),
                       (5, [1, 2, 1, 1, 2])):
        cx = O.Complex.kuhn(dim, shape)
        assert fq.kuhn_counts(dim, shape) == [cx.nsimplices(j) for j in range(dim + 1)]
        for j in range(dim + 1):
            got = fq.kuhn_cell_faces_host(dim, shape, j).astype(np.int64)
            assert np.array_equal(got, cx.cell_faces(j)), (dim, shape, j)


def test_kuhn_counts_polynomials(fq):
    # V=(N+1)^3, E=7N^3+9N^2+3N, F=12N^3+6N^2, C=6N^3
    for N in (1, 2, 5, 128):
        assert fq.kuhn_counts(3, N) == [(N + 1) ** 3, 7 * N ** 3 + 9 * N ** 2 + 3 * N, 12 * N ** 3 + 6 * N ** 2, 6 * N ** 3]
    for N in (1, 3, 16):
        assert fq.kuhn_counts(4, N) == [(N + 1) ** 4, 15 * N ** 4 + 28 * N ** 3 + 18 * N ** 2 + 4 * N,
                                        50 * N ** 4 + 48 * N ** 3 + 12 * N ** 2, 60 * N ** 4 + 24 * N ** 3, 24 * N ** 4]


def test_interface_mirrors_reference_grades(fq):
    # operators.rs:195-200: a differentiated side sits one grade below
    W = fq.WhitneyPairing
    assert (W.mass(3, 1).test_grade(), W.mass(3, 1).trial_grade()) == (1, 1)
    assert (W.dif_trial(3, 1).test_grade(), W.dif_trial(3, 1).trial_grade()) == (1, 0)
    assert (W.dif_test(3, 1).test_grade(), W.dif_test(3, 1).trial_grade()) == (0, 1)
    assert (W.dif_both(3, 2).test_grade(), W.dif_both(3, 2).trial_grade()) == (1, 1)
    assert W.dif_test(3, 1).element_shape() == (4, 6)
    assert fq.ScalarLumpedMass(3).element_shape() == (4, 4)


def test_no_cpu_fallback(fq):
    from formoniq_b200 import _lib

    if _lib.lib().fq_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(fq.FormoniqError):
        fq.Context(0)


def test_product_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "formoniq_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")):
                text = open(os.path.join(dp, f), errors="replace").read()
                assert "fq_oracle" not in text and "liboracle" not in text and "import oracle" not in text and \
                    "from oracle" not in text, os.path.join(dp, f)


def test_eigen_seed_vectors_match_the_reference_probe():
    # eigen.rs:259-268: the vectorised splitmix64 fill of formoniq_b200.eigen equals the oracle's scalar restatement
    import numpy as np

    from formoniq_b200.eigen import MASK, pseudo_random
    from oracle import oracle as O

    for seed in (0, 1, 7, MASK):
        v = pseudo_random(seed, 64)
        assert all(v[i] == O.pseudo_random(seed, i) for i in range(64))
        assert np.all(np.abs(v) <= 1.0)


def _build_abi_smoke(tmp_path):
    from formoniq_b200 import build as B

    lib = B.build()
    exe = tmp_path / "abi_smoke"
    subprocess.check_call(["/usr/bin/gcc", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c", "abi_smoke.c"), "-o", str(exe), "-L", os.path.dirname(lib),
                           "-lformoniq_b200", "-Wl,-rpath," + os.path.dirname(lib)])
    return exe


def test_header_compiles_as_c_and_reports_the_missing_device(tmp_path):
    # include/formoniq_b200.h exercised by a C compiler (not ctypes): without a device every compute entry point must
    # fail with FQ_ERR_CUDA — there is no CPU fallback to route through
    import torch

    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    out = subprocess.run([str(_build_abi_smoke(tmp_path))], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "ok nogpu", out.stdout + out.stderr


@pytest.mark.gpu
def test_c_program_assembles_through_the_abi(tmp_path):
    out = subprocess.run([str(_build_abi_smoke(tmp_path))], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "ok gpu", out.stdout + out.stderr
