"""The boundary is a C ABI: a plain C host (examples/host_demo.c -- no Python, no torch) compiles against include/p2e.h,
links libp2e.so, uses the host-only entry points (tape layout, generator-stream loader) and, without a device, gets the
library's loud failure instead of any CPU path; on a GPU box the same binary runs a witness batch."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_and_run(tmp_path, name="host_demo"):
    import plonky2_ecdsa_b200 as pkg
    if not os.path.exists(pkg.LIB_PATH):
        from plonky2_ecdsa_b200.build import build
        build()
    exe = str(tmp_path / name)
    libdir = os.path.dirname(pkg.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-pthread", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", name + ".c"), "-L", libdir, "-lp2e", "-Wl,-rpath," + libdir, "-o", exe])
    return subprocess.run([exe], capture_output=True, text=True, timeout=600)


def test_c_host_links_and_fails_loudly_without_a_device(tmp_path):
    import torch
    r = _build_and_run(tmp_path)
    assert "4642 generator instances, 82615 witness slots" in r.stdout
    assert "2 generators, 2 levels, 69 slots, 18 input slots" in r.stdout
    if torch.cuda.is_available():
        assert r.returncode == 0 and "1000 mulmod witnesses ok" in r.stdout
    else:
        assert r.returncode == 2 and "no device" in r.stderr


@pytest.mark.gpu
def test_c_host_runs_a_witness_batch(tmp_path):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device in this container")
    r = _build_and_run(tmp_path)
    assert r.returncode == 0, r.stderr
    assert "1000 mulmod witnesses ok" in r.stdout


def test_c_multi_device_host_builds_and_fails_loudly_without_a_device(tmp_path):
    import torch
    r = _build_and_run(tmp_path, "host_multi")
    if not torch.cuda.is_available():
        assert r.returncode == 2 and "no device" in r.stderr


@pytest.mark.gpu
def test_c_host_drives_every_device_from_one_process(tmp_path):
    """examples/host_multi.c: one process, one host thread per GPU (p2e_init_devices / p2e_set_device), shards of one ECDSA
    batch; then two threads issuing host-pointer calls on the same device concurrently.  With one GPU visible the shards
    share device 0 (the threading contract is still exercised); with 2+ GPUs the devices run side by side."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device in this container")
    r = _build_and_run(tmp_path, "host_multi")
    assert r.returncode == 0, r.stderr + r.stdout
    assert "identical to device 0" in r.stdout and "all identical" in r.stdout
