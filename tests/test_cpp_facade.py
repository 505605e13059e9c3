"""The C++ facade (include/kmerseek_b200.hpp) compiles against the C ABI and behaves like the reference's
builder / create_protein_signature / store_signatures (src/rust/index.rs:1395-1441, 3021-3036)."""
import os
import subprocess

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def demo(tmp_path_factory):
    out = tmp_path_factory.mktemp("cpp") / "facade_demo"
    libdir = os.path.join(ROOT, "kmerseek_b200")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "facade_demo.cpp"), "-L", libdir, "-lkmerseek_b200",
           f"-Wl,-rpath,{libdir}", "-o", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return str(out)


def test_facade_compiles_and_builder_errors(demo):
    r = subprocess.run([demo, "builder"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "builder ok" in r.stdout


@pytest.mark.gpu
def test_facade_on_gpu(demo, golden_rust):
    r = subprocess.run([demo, "gpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = dict((l.split()[0], l.split()[1:]) for l in r.stdout.strip().splitlines())
    assert lines["protein"][:2] == ["17", "17"] and lines["protein"][2] == "7641839ad508ab8"  # index.rs:1417-1439
    assert lines["dayhoff"][:2] == ["17", "17"]
    assert lines["hp"][:2] == ["14", "14"]  # index.rs:1519-1541
