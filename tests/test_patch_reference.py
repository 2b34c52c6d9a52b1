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


# ---------------------------------------------------------------------------------------------
# Gate / LQE patches on the real reference classes (transformer.py:222-235, 263-288)
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def ref_transformer():
    spec = importlib.util.spec_from_file_location(
        "make_golden_blocks_patch_test", os.path.join(os.path.dirname(__file__), "golden", "make_golden_blocks.py"))
    mgb = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mgb)
    return mgb.load_reference_transformer()


def test_gate_patch_on_reference_class(ref_transformer):
    from detrpose_b200.gate import install_gate, uninstall_gate, Gate
    T = ref_transformer
    m = T.Gate(128)
    original = T.Gate.forward
    x1, x2 = torch.randn(2, 3, 128), torch.randn(2, 3, 128)
    want = m(x1, x2)
    install_gate(T)
    try:
        assert T.Gate.forward is not original
        assert torch.equal(m(x1, x2), want)                  # CPU tensors: the reference's own forward
    finally:
        uninstall_gate(T)
    assert T.Gate.forward is original
    # the drop-in class loads the reference's state_dict and starts from the reference's initialisation
    ours = Gate(128)
    assert sorted(ours.state_dict()) == sorted(m.state_dict())
    for k, v in m.state_dict().items():
        assert torch.equal(ours.state_dict()[k], v), k
    ours.load_state_dict(m.state_dict())


def test_lqe_patch_on_reference_class(ref_transformer):
    from detrpose_b200.lqe import install_lqe, uninstall_lqe, LQE
    T = ref_transformer
    m = T.LQE(4, 32, 2, 17)
    original = T.LQE.forward
    scores, poses, feat = torch.randn(1, 3, 1), torch.rand(1, 3, 34), torch.randn(1, 128, 6, 5)
    want = m(scores, poses, feat)
    install_lqe(T)
    try:
        assert torch.equal(m(scores, poses, feat), want)     # CPU tensors: the reference's own forward
    finally:
        uninstall_lqe(T)
    assert T.LQE.forward is original
    ours = LQE(4, 32, 2, 17)
    assert sorted(ours.state_dict()) == sorted(m.state_dict())
    assert [tuple(v.shape) for _, v in sorted(ours.state_dict().items())] == \
           [tuple(v.shape) for _, v in sorted(m.state_dict().items())]
    ours.load_state_dict(m.state_dict())
