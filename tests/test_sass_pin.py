"""The row-major search kernel of the headline layout is the one profiles/r1_v10_* measured: its SASS is
pinned (profiles/r1_v10_sass_pin.json).  Later work (column-major evaluator, schedules, host emulation
hooks) must not perturb it silently — a deliberate change re-pins it together with new measurements."""
import hashlib
import json
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_row_major_headline_kernel_is_the_measured_one():
    pin = json.load(open(os.path.join(ROOT, "profiles", "r1_v10_sass_pin.json")))
    obj = os.path.join(ROOT, pin["object"])
    if not os.path.exists(obj) or not shutil.which("cuobjdump"):
        pytest.skip("object files of the in-tree build or cuobjdump not available")
    nvcc = subprocess.run(["nvcc", "--version"], capture_output=True, text=True).stdout if shutil.which("nvcc") else ""
    if "12.9" not in nvcc:
        pytest.skip("pinned for the nvcc 12.9 of this image")
    out = subprocess.run(["cuobjdump", "-sass", "-fun", pin["mangled"], obj], capture_output=True, text=True, check=True).stdout
    ins = [m.group(1).strip() for m in (re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(.*?);", l) for l in out.splitlines()) if m]
    assert len(ins) == pin["instructions"]
    assert hashlib.sha256("\n".join(ins).encode()).hexdigest() == pin["sha256_of_sass_text"]
