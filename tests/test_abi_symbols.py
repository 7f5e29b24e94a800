"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports exactly the
symbols include/rumdeed_b200.h declares, and refuses to compute without a GPU (no fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

import rumdeed_b200 as rb
from rumdeed_b200.api import EXPORTS, LIB_PATH

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "rumdeed_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(rb2_[a-z0-9_]+)\s*\(", txt)))


def test_header_matches_binding_table():
    assert header_symbols() == sorted(EXPORTS)


def test_library_exports_every_header_symbol():
    assert os.path.exists(LIB_PATH), "run __graft_entry__.build() first"
    lib = rb.load_library()
    for s in header_symbols():
        assert hasattr(lib, s), s
    nm = subprocess.run(["nm", "-D", "--defined-only", LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (rb2_[a-z0-9_]+)", nm))
    assert set(header_symbols()) <= exported
    # nothing from the oracle is linked into the product
    assert "orc_" not in nm


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "rumdeed_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), os.path.join(dirpath, f)


def test_struct_sizes_match_header():
    # compile a tiny C program against the header and compare sizeof with the ctypes mirrors
    code = r'''
#include <stdio.h>
#include "rumdeed_b200.h"
#include "rumdeed_host.h"
int main(void){ printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(rb2_config), sizeof(rb2_counts), sizeof(rb2_event),
                       sizeof(rb2_step_result), sizeof(rb2_mh_config), sizeof(rb2_collision_config), sizeof(rb2_recomb_record),
                       sizeof(rb2_ionization_record), sizeof(rb2_collision_result), sizeof(rh_state), sizeof(rh_setup)); return 0; }
'''
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        src = os.path.join(td, "s.c")
        open(src, "w").write(code)
        exe = os.path.join(td, "s")
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()
    sizes = [int(x) for x in out]
    from rumdeed_b200 import api, host_api
    assert sizes == [C.sizeof(rb.Config), C.sizeof(rb.Counts), C.sizeof(rb.Event), C.sizeof(rb.StepResult), C.sizeof(api.MhConfig),
                     C.sizeof(api.CollisionConfig), C.sizeof(api.RecombRecord), C.sizeof(api.IonizationRecord),
                     C.sizeof(api.CollisionResult), C.sizeof(host_api.State), C.sizeof(host_api.Setup)]


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = rb.load_library()
    assert lib.rb2_device_available() == 0
    cfg = rb.planar_config(2.0, 1e-7, (1e-7, 1e-7, 1e-7), 1e-16, True, 1, capacity=16)
    with pytest.raises(rb.Rb2Error) as e:
        rb.HotPath(cfg)
    assert "no CPU fallback" in str(e.value) or "rb2 error -2" in str(e.value)
    # uninitialised calls report RB2_ERR_NOT_INIT instead of computing anything
    assert lib.rb2_accel_only() == -1
    assert lib.rb2_field_batch(1, None, None) == -1


def test_bench_reference_arm_runs_on_the_cpu():
    """`bench.py --impl reference` (the CPU restatement of the reference's pair loop on the host cores) needs no GPU and
    none of the product: one JSON line with the contract's keys, the thread count stated, the same row sample rule."""
    import json
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--particles", "3000", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600, env={**os.environ, "OMP_NUM_THREADS": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "pair_interactions_per_s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["e2e"]["value"] == line["value"] and line["gpu_launches"] == 0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and "100%" in cb["sample"].replace(" ", "")  # every row at this size
    # the thread count is set explicitly: OMP_NUM_THREADS=1 (what torchrun exports) does not pin the arm to one core
    assert cb["cores"] == len(os.sched_getaffinity(0))
