"""Compiles and runs the C++ host layer (include/trexb200.hpp) against libtrexb200.so."""
import os
import subprocess

import pytest

from conftest import ROOT


def _build(tmp_path):
    exe = str(tmp_path / "test_shim")
    lib_dir = os.path.join(ROOT, "trex_b200")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_shim.cpp"),
                    "-o", exe, "-L", lib_dir, "-ltrexb200", f"-Wl,-rpath,{lib_dir}"], check=True)
    return exe


def test_cpp_shim_compiles(tmp_path):
    import __graft_entry__ as g
    g.build()
    _build(tmp_path)


@pytest.mark.gpu
def test_cpp_shim_runs(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout + r.stderr


def _build_plugin(tmp_path):
    exe = str(tmp_path / "test_plugin")
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_plugin.c"),
                    "-o", exe, "-ldl"], check=True)
    return exe


def test_plugin_registers_through_dlopen(tmp_path):
    """dlopen + the one register symbol + the pure-C back-end table (SURVEY s8b A''); compiled as C.  Runs without a GPU: the table is
    complete and init refuses to start without a device."""
    import __graft_entry__ as g
    g.build()
    exe = _build_plugin(tmp_path)
    r = subprocess.run([exe, os.path.join(ROOT, "trex_b200", "libtrexb200.so")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout + r.stderr


@pytest.mark.gpu
def test_plugin_runs_detection_and_identification(tmp_path):
    exe = _build_plugin(tmp_path)
    r = subprocess.run([exe, os.path.join(ROOT, "trex_b200", "libtrexb200.so"), "gpu"], capture_output=True, text=True, timeout=180)
    assert r.returncode == 0 and r.stdout.startswith("OK plug-in"), r.stdout + r.stderr
