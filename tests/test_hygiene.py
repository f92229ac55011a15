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
        # the CPU legs: the port's rate, the reference arm, and the exact (HiGHS) solve timed beside them
        assert name in ("cpu_port_rate", "reference_arm", "exact_solve"), name


def test_host_emulation_stays_out_of_the_product():
    """tests/emu compiles the device headers for the host (KAO_HOST_EMU) as a checker; the product's
    build never defines that macro and libkao.so carries none of the emulation's symbols."""
    import subprocess

    mk = open(os.path.join(PKG, "csrc", "Makefile")).read()
    assert "KAO_HOST_EMU" not in mk and "tests/" not in mk and "emu" not in mk
    for base, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".hpp")):
                text = open(os.path.join(base, f), errors="ignore").read()
                assert "kao_emu" not in text and "warp_emu" not in text, f
    so = os.path.join(PKG, "libkao.so")
    if os.path.exists(so):
        syms = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True).stdout
        assert "kao_emu" not in syms and "kao_solve" in syms


def test_bench_clock_sampler_says_so_when_there_is_no_nvml():
    """bench.py samples clocks through NVML during the timed region; without a driver it must neither hang nor
    invent numbers: no samples, an `error` field."""
    import importlib.util
    import time

    spec = importlib.util.spec_from_file_location("kao_bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    s = bench.ClockSampler(0)
    t0 = time.time()
    s.start()
    s.mark_begin()
    time.sleep(0.03)
    s.mark_end()
    out = s.stop()
    assert time.time() - t0 < 15.0
    if out["samples"] == 0:                      # this container: no GPU, no libnvidia-ml
        assert out["sm_mhz"] is None and out["reasons"] == [] and "error" in out
    else:                                        # a GPU box: real samples, all of them inside or around the region
        assert out["sm_mhz"] and out["sm_max_mhz"] and out["samples_in_timed_region"] <= out["samples"]
