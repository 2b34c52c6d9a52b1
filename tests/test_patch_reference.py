"""`detrpose_b200.patch` against the real reference module (build container only: skipped where
/root/reference is absent).  No kernels run here: the checks are about routing and about leaving the
reference class untouched."""
import importlib.util
import os

import pytest
import torch

import detrpose_b200 as dp

REF_FILE = "/root/reference/src/models/detrpose/ms_deform_attn.py"
pytestmark = pytest.mark.skipif(not os.path.exists(REF_FILE), reason="reference checkout not present")


@pytest.fixture()
def ref():
    spec = importlib.util.spec_from_file_location("ref_msda_patch_test", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _inputs(n_levels=2, n_points=4):
    torch.manual_seed(0)
    shapes = [[4, 6], [2, 3]][:n_levels]
    S = sum(h * w for h, w in shapes)
    mem = torch.randn(2, S, 64)
    value = mem.unflatten(2, (8, -1)).permute(0, 2, 3, 1).flatten(0, 1).split([h * w for h, w in shapes], dim=-1)
    return torch.randn(2, 10, 64), torch.rand(2, 2, 1, 5, 2), value, shapes


def test_install_is_reversible_and_idempotent(ref):
    original = ref.ms_deform_attn_core_pytorch
    dp.patch.install(ref)
    dp.patch.install(ref)
    assert ref.ms_deform_attn_core_pytorch is not original
    dp.patch.uninstall(ref)
    assert ref.ms_deform_attn_core_pytorch is original


def test_baseline_configuration_has_no_cpu_path(ref):
    q, refp, value, shapes = _inputs()
    mod = ref.MSDeformAttn(d_model=64, n_levels=2, n_heads=8, n_points=4)
    dp.patch.install(ref)
    try:
        with pytest.raises(RuntimeError, match="no CPU path"):
            mod(q, refp, value, shapes)
    finally:
        dp.patch.uninstall(ref)


def test_cpu_can_be_handed_back_to_the_reference_explicitly(ref):
    q, refp, value, shapes = _inputs()
    mod = ref.MSDeformAttn(d_model=64, n_levels=2, n_heads=8, n_points=4)
    expected = mod(q, refp, value, shapes)
    dp.patch.install(ref, cpu_to_reference=True)
    try:
        assert torch.equal(mod(q, refp, value, shapes), expected)
    finally:
        dp.patch.uninstall(ref)


def test_optional_branches_are_routed_to_the_reference(ref):
    q, refp, value, shapes = _inputs()
    mod = ref.MSDeformAttn(d_model=64, n_levels=2, n_heads=8, n_points=4, use_modulation=True)
    expected = mod(q, refp, value, shapes)
    dp.patch.install(ref)
    try:
        assert torch.equal(mod(q, refp, value, shapes), expected)      # modulation: reference's own code
    finally:
        dp.patch.uninstall(ref)


def test_drop_in_module_interface_equals_reference_class(ref):
    theirs = ref.MSDeformAttn(d_model=128, n_levels=3, n_heads=8, n_points=4)
    ours = dp.MSDeformAttn(d_model=128, n_levels=3, n_heads=8, n_points=4)
    assert [k for k, _ in ours.named_parameters()] == [k for k, _ in theirs.named_parameters()]
    for (_, a), (_, b) in zip(ours.named_parameters(), theirs.named_parameters()):
        assert a.shape == b.shape and torch.equal(a, b)
    ours.load_state_dict(theirs.state_dict())
