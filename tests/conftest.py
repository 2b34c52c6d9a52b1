"""Shared fixtures.  `-m gpu` tests need a B200 and call the kernels through the C ABI;
everything else runs on CPU (oracle vs golden vectors, host logic, symbol checks)."""
import glob
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


# (forward variant, backward variant) pairs that are ALLOWED to refuse a shape (the library fails loudly,
# the test reports a skip).  The default pair (1, 1) = lean forward + gather backward and the flat pair are
# not in the list: a refusal from them is a failure, so a regression that makes the default kernels reject a
# model shape cannot hide behind a skip.  Only the opt-in TMA-staged forward has shape limits by design
# (it needs two or more levels whose coarse levels fit shared memory, bf16 / fp32 rows of 64 or 128 bytes).
MAY_REFUSE = {(2, 1): "staged forward does not support this shape",
              (3, 2): "staged forward does not support this shape"}


@pytest.hookimpl(hookwrapper=True)
def pytest_runtest_call(item):
    """A kernel variant forced by the `bwd_variant` fixture may refuse shapes it does not support -- but only
    the pairs, and with the message, listed in MAY_REFUSE; anything else propagates as a failure."""
    outcome = yield
    if outcome.excinfo is None or "bwd_variant" not in getattr(item, "fixturenames", ()):
        return
    callspec = getattr(item, "callspec", None)
    pair = tuple(callspec.params.get("bwd_variant", ())) if callspec is not None else ()
    allowed = MAY_REFUSE.get(pair)
    if allowed is not None and allowed in str(outcome.excinfo[1]):
        outcome.force_exception(pytest.skip.Exception(f"variant {pair}: {allowed}"))


def core_case_names():
    return sorted(os.path.basename(p)[len("core_"):-len(".npz")]
                  for p in glob.glob(os.path.join(GOLDEN_DIR, "core_*.npz")))


def load_core_case(name):
    z = np.load(os.path.join(GOLDEN_DIR, f"core_{name}.npz"))
    case = {k: z[k] for k in z.files}
    case["shapes"] = tuple((int(h), int(w)) for h, w in case["shapes"])
    case["n_heads"] = int(case["n_heads"])
    return case


def load_module_case():
    z = np.load(os.path.join(GOLDEN_DIR, "module_small.npz"))
    return {k: z[k] for k in z.files}


def value_list_from_memory(memory, n_heads, shapes):
    """The reference caller's value construction (transformer.py:1285-1286) on a torch tensor."""
    sizes = [h * w for h, w in shapes]
    v = memory.unflatten(2, (n_heads, -1)).permute(0, 2, 3, 1).flatten(0, 1)
    return list(v.split(sizes, dim=-1))


def rel_err(a, b):
    """max |a-b| relative to max |b| (the tolerance form BASELINE.json's north_star states)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = max(float(np.abs(b).max()), 1e-30)
    return float(np.abs(a - b).max()) / denom
