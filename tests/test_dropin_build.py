"""Drop-in boundary: the reference's only in-tree caller, demo/main.cpp, must compile UNCHANGED
against include/atomorph/*.h and link against libatomorph_b200.so.  Needs
the reference tree, so it only runs in the build container."""
import os
import shutil
import subprocess

import pytest

from atomorph_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
def test_demo_compiles_and_links_unchanged(tmp_path):
    demo = tmp_path / "demo"
    demo.mkdir()
    for f in ("main.cpp", "main.h", "options.h", "lodepng.cpp", "lodepng.h"):
        shutil.copy(os.path.join(REF, "demo", f), demo / f)
    # demo/main.h includes "../atomorph.h": that path now resolves to the drop-in header
    (tmp_path / "atomorph.h").write_text('#include "atomorph/atomorph.h"\n')
    exe = demo / "atomorph_demo"
    cmd = ["g++", "-std=c++11", "-O1", "-w", "-I", os.path.join(ROOT, "include"), "main.cpp", "lodepng.cpp", "-o", str(exe),
           "-L", os.path.dirname(_lib.LIB_PATH), "-latomorph_b200", "-Wl,-rpath," + os.path.dirname(_lib.LIB_PATH), "-pthread"]
    r = subprocess.run(cmd, cwd=demo, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([str(exe), "--version"], capture_output=True, text=True, timeout=30)
    assert r.returncode == 0
