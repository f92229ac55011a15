"""The product never touches the oracle: no file of the package (Python or C++/CUDA) imports,
includes, links or names anything under oracle/ — there is no CPU fallback to route through."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "kafka_assignment_optimizer_b200")


def test_product_sources_do_not_reference_the_oracle():
    offenders = []
    for base, _, files in os.walk(PKG):
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")) and f != "Makefile":
                continue
            text = open(os.path.join(base, f), errors="ignore").read()
            if re.search(r"\b(from|import)\s+oracle\b|oracle/|kao_ref|libkao_ref", text):
                offenders.append(os.path.relpath(os.path.join(base, f), ROOT))
    assert offenders == []
    hdr = open(os.path.join(ROOT, "include", "kao.h")).read()
    assert "oracle" not in hdr


def test_bench_uses_the_oracle_only_in_its_cpu_legs():
    src = open(os.path.join(ROOT, "bench.py")).read()
    for mt in re.finditer(r"from oracle import", src):
        fn = src.rfind("\ndef ", 0, mt.start())
        name = re.match(r"\ndef (\w+)", src[fn:]).group(1)
        assert name in ("cpu_port_rate", "reference_arm"), name
