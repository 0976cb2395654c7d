import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def load_fixture(name):
    return torch.load(os.path.join(GOLDEN, f"{name}.pt"), weights_only=False)


def grid_of(fix):
    g = fix["grid_size"]
    return [g] * 3 if isinstance(g, int) else list(g)


def oracle_scene(fix, **kw):
    """oracle.nmf_oracle.Scene of a golden fixture (test side only)."""
    from oracle import nmf_oracle
    model = "microfacet" if fix["model"] == "microfacet_tensorf2" else "plain"
    return nmf_oracle.Scene(fix["state"], fix["aabb"], fix["near_far"], grid_of(fix),
                            alpha_volume=fix["alpha_volume"].float(), model=model, **kw)


def device_scene(fix, device, **kw):
    from nmf_b200.scene import DeviceScene
    model = "microfacet" if fix["model"] == "microfacet_tensorf2" else "plain"
    return DeviceScene(fix["state"], fix["aabb"], fix["near_far"], grid_of(fix),
                       alpha_volume=fix["alpha_volume"], device=device, model=model, **kw)


@pytest.fixture(scope="session")
def hostcheck():
    """libnmf_hostcheck.so: the kernels' per-element math compiled for the host (tests/hostcheck)."""
    import ctypes
    d = os.path.join(ROOT, "tests", "hostcheck")
    so = os.path.join(d, "libnmf_hostcheck.so")
    srcs = [os.path.join(d, "hostcheck.cpp"), os.path.join(ROOT, "nmf_b200", "csrc", "nmf_math.cuh"),
            os.path.join(ROOT, "nmf_b200", "csrc", "nmf_field.cuh"), os.path.join(ROOT, "nmf_b200", "csrc", "nmf_train.cuh"),
            os.path.join(ROOT, "nmf_b200", "csrc", "nmf_microfacet_bwd.cuh"),
            os.path.join(ROOT, "include", "nmf_b200.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, srcs[0]])
    return ctypes.CDLL(so)
