"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/kao.h
declares (no compute without a GPU), the Python builder mirrors the oracle's defaults, the JSON
codec keeps the reference's shapes (README.md:52-63 in, :67-78 out)."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

import kafka_assignment_optimizer_b200 as kao
from kafka_assignment_optimizer_b200 import optimizer as kopt, problem as kprob
from oracle import model as m

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g

    if not os.path.exists(kopt._LIB_PATH):
        g.build()
    return kopt.load_library()


def test_abi_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "kao.h")).read()
    names = set(re.findall(r"\b(kao_[a-z0-9_]+)\s*\(", hdr))
    assert {"kao_solve", "kao_eval", "kao_create", "kao_search", "kao_round_launch", "kao_round_apply",
            "kao_candidate_keys", "kao_profile_rounds", "kao_p2p_export", "kao_p2p_connect",
            "kao_search_sharded", "kao_search_sharded_delta", "kao_search_delta", "kao_set_patience", "kao_set_evaluator", "kao_get_evaluator", "kao_set_schedule", "kao_last_rounds", "kao_candidate_keys_delta", "kao_key_obj_bits", "kao_version", "kao_last_error"} <= names
    for n in sorted(names):
        assert hasattr(lib, n), n
    assert lib.kao_version() == 0x00020000


def test_struct_layout_matches_header(tmp_path):
    """The ctypes mirrors of kao_problem / kao_options / kao_result have the size and field offsets the C
    compiler gives the structs of include/kao.h (asked of gcc, not hard-coded)."""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    mirrors = {"kao_problem": kopt._KaoProblem, "kao_options": kopt._KaoOptions, "kao_result": kopt._KaoResult}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "kao.h"', 'int main(void) {']
    for cname, cls in mirrors.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ['return 0; }']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(root, "include"), "-o", str(exe), str(src)])
    want = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, cls in mirrors.items():
        assert ctypes.sizeof(cls) == int(want[cname]), cname
        for fname, _ in cls._fields_:
            assert getattr(cls, fname).offset == int(want["%s.%s" % (cname, fname)]), (cname, fname)


def test_no_gpu_means_loud_failure(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(kao.KaoError, match="no CUDA device"):
        kopt.solve(kao.synthetic_problem(16, 8, 2, 2), rounds=1, round_size=16)


def test_builder_mirrors_oracle_defaults():
    for args in [(256, 32, 4, 3, 2), (100, 20, 3, 2, 1), (64, 12, 5, 3, 0)]:
        a, b = kao.synthetic_problem(*args), m.synthetic_problem(*args)
        for f in ("P", "B", "R", "RF", "ppr_lo", "ppr_hi"):
            assert getattr(a, f) == getattr(b, f)
        for f in ("rack_of", "wF", "wL", "rep_lo", "rep_hi", "ldr_lo", "ldr_hi", "rack_lo", "rack_hi", "cur"):
            assert np.array_equal(getattr(a, f), getattr(b, f)), f
    a, b = kao.synthetic_problem(64, 16, 4, 3, 0, 0.1, 9), m.synthetic_problem(64, 16, 4, 3, 0, 0.1, 9)
    assert np.array_equal(a.cur, b.cur)


README_CURRENT = """{"version":1,"partitions":[
    {"topic":"x.y.z.t","partition":0,"replicas":[7,18]}, {"topic":"x.y.z.t","partition":1,"replicas":[8,19]},
    {"topic":"x.y.z.t","partition":2,"replicas":[9,10]}, {"topic":"x.y.z.t","partition":3,"replicas":[0,11]},
    {"topic":"x.y.z.t","partition":4,"replicas":[1,12]}, {"topic":"x.y.z.t","partition":5,"replicas":[2,13]},
    {"topic":"x.y.z.t","partition":6,"replicas":[3,14]}, {"topic":"x.y.z.t","partition":7,"replicas":[4,15]},
    {"topic":"x.y.z.t","partition":8,"replicas":[5,16]}, {"topic":"x.y.z.t","partition":9,"replicas":[6,17]}]}"""


def test_json_codec_readme_shapes():
    rows, topics = kprob.parse_assignment_json(README_CURRENT)           # README.md:52-63
    assert rows[1] == [8, 19] and topics[9] == ("x.y.z.t", 9)
    brokers = kprob.parse_broker_list("0,1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,16,17,18")   # README.md:48
    racks = kprob.parse_rack_map(",".join("%d:%s" % (b, "b" if b % 2 else "a") for b in range(20)))
    pb = kao.build_problem(rows, brokers, racks, 2, topics)
    ref = m.readme_problem()
    assert np.array_equal(pb.cur, ref.cur) and np.array_equal(pb.wL, ref.wL) and np.array_equal(pb.rack_of, ref.rack_of)
    assert pb.cur[1].tolist() == [8, -1]                                  # broker 19 is not in the target list
    doc = kprob.reassignment_json(pb, np.array([[7, 18], [8, 1]] + [[0, 1]] * 8))
    assert doc["version"] == 1 and doc["partitions"][1] == {"topic": "x.y.z.t", "partition": 1, "replicas": [8, 1]}
    json.dumps(doc)


def test_jni_shim_type_checks_against_the_abi():
    """java/kao_jni.c cannot be built here (no JDK): type-check it against include/kao.h with a stub
    <jni.h> that declares the JNI entries it uses with the specification's signatures."""
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-fsyntax-only",
                        "-I", os.path.join(root, "tests", "jni_stub"), "-I", os.path.join(root, "include"),
                        os.path.join(root, "java", "kao_jni.c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_rounds_are_bounded_before_anything_is_allocated(lib):
    """`rounds` sized host vectors unchecked (std::bad_alloc across extern "C"; -1 wraps to 4 billion
    through c_uint32).  KAO_MAX_ROUNDS is checked first, and every entry point runs inside a catch-all guard."""
    pb = kao.synthetic_problem(16, 8, 2, 2)
    for rounds in ((1 << 20) + 1, -1):
        with pytest.raises(kao.KaoError, match="KAO_MAX_ROUNDS"):
            kopt.solve(pb, rounds=rounds, round_size=16)
    with pytest.raises(kao.KaoError, match="round_size"):
        kopt.solve(pb, rounds=1, round_size=1 << 25)
    assert kopt.key_obj_bits(pb) == (16 * 2 * 4).bit_length()       # needs no GPU
