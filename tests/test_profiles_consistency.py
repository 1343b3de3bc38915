"""CPU: the committed ncu evidence that bench.py quotes (profiles/dominant_kernel_traffic.json -> roofline.traffic) must describe the
kernels of the library as built now: every entry carries the register count of the captured kernel, and it has to equal what
`cuobjdump -res-usage` reports for the current build (the traffic entry silently described an older 64-register build)."""
import json
import os
import re
import shutil
import subprocess

import pytest

from conftest import ROOT


def test_traffic_entries_match_the_built_library():
    path = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    if not os.path.isfile(path) or shutil.which("cuobjdump") is None:
        pytest.skip("no traffic file / no cuobjdump")
    from sola_b200 import _build
    res = subprocess.run(["cuobjdump", "-res-usage", _build.build()], capture_output=True, text=True).stdout
    norm = lambda n: re.sub(r"\bfalse\b", "0", re.sub(r"\btrue\b", "1", n))      # c++filt prints bool template arguments as words, ncu as 0 / 1
    regs = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+)", res):
        dn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        regs[norm(re.sub(r"^void\s+", "", dn).split("(")[0].replace("sola::", ""))] = int(m.group(2))
    entries = json.load(open(path))
    checked = 0
    for key, e in entries.items():
        if "registers" not in e or "kernel" not in e:
            continue
        name = norm(re.sub(r"^void\s+", "", e["kernel"]).replace("sola::", ""))
        assert name in regs, f"{key}: kernel {name} is not in the built library"
        assert regs[name] == e["registers"], f"{key}: captured with {e['registers']} registers, the library now has {regs[name]} — re-capture (tools/gpu_profile_r2.sh)"
        checked += 1
    assert checked >= 1
