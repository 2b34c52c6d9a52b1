"""Host-side logic that needs no GPU: module interface parity with the reference (constructor,
parameter names, initialisation, state_dict), sharding arithmetic, byte accounting, and the
refusal to run on CPU tensors."""
import inspect

import numpy as np
import pytest
import torch

import detrpose_b200 as dp
from detrpose_b200 import shard, synthetic
from conftest import load_module_case


def test_constructor_signature_matches_reference():
    # ms_deform_attn.py:197-203
    params = inspect.signature(dp.MSDeformAttn.__init__).parameters
    expected = dict(d_model=256, n_levels=4, n_heads=8, n_points=4, use_4D_normalizer=False,
                    use_modulation=False, use_region_sampling=False, region_kernel_size=1,
                    use_global_context=False, use_grouped_offsets=False, num_groups=1,
                    use_grid_attention=False, grid_num_points=16, use_grid_offsets=False,
                    use_grid_fusion=True, is_energy=False)
    assert list(params)[1:] == list(expected)
    for k, v in expected.items():
        assert params[k].default == v
    fwd = list(inspect.signature(dp.MSDeformAttn.forward).parameters)
    assert fwd == ["self", "query", "reference_points", "value", "input_spatial_shapes"]


def test_parameter_names_shapes_and_init_match_reference():
    m = load_module_case()
    d_model, L, H, P = [int(v) for v in m["hyper"]]
    mod = dp.MSDeformAttn(d_model=d_model, n_levels=L, n_heads=H, n_points=P)
    sd = mod.state_dict()
    ref_keys = sorted(k[len("init."):] for k in m if k.startswith("init."))
    assert sorted(sd) == ref_keys
    for k in ref_keys:
        assert tuple(sd[k].shape) == m[f"init.{k}"].shape
        assert np.array_equal(sd[k].numpy(), m[f"init.{k}"]), k
    # n_points % 4 != 0 -> zero offset bias (DETRPose-N, ms_deform_attn.py:311-312)
    mod_n = dp.MSDeformAttn(d_model=128, n_levels=2, n_heads=8, n_points=6)
    for k, v in mod_n.state_dict().items():
        assert np.array_equal(v.numpy(), m[f"init_n.{k}"]), k


def test_reference_state_dict_loads():
    m = load_module_case()
    d_model, L, H, P = [int(v) for v in m["hyper"]]
    mod = dp.MSDeformAttn(d_model=d_model, n_levels=L, n_heads=H, n_points=P)
    state = {k[len("param."):]: torch.from_numpy(v) for k, v in m.items() if k.startswith("param.")}
    missing, unexpected = mod.load_state_dict(state, strict=True)
    assert not missing and not unexpected


@pytest.mark.parametrize("flag", ["use_modulation", "use_region_sampling", "use_global_context",
                                  "use_grouped_offsets", "use_grid_attention", "is_energy"])
def test_optional_branches_are_refused(flag):
    with pytest.raises(NotImplementedError, match=flag):
        dp.MSDeformAttn(**{flag: True})


def test_bad_head_split_raises_like_reference():
    with pytest.raises(ValueError, match="divisible"):
        dp.MSDeformAttn(d_model=100, n_heads=8)


def test_cpu_tensors_are_refused():
    loc = torch.rand(1, 2, 2, 1, 2, 2)
    att = torch.full((1, 2, 2, 1, 2), 0.5)
    val = [torch.randn(2, 8, 16)]
    with pytest.raises(RuntimeError, match="no CPU path"):
        dp.ms_deform_attn_core(val, [(4, 4)], loc, att)


def test_core_argument_validation():
    loc = torch.rand(1, 2, 2, 1, 2, 2)
    with pytest.raises((RuntimeError, ValueError)):
        dp.ms_deform_attn_core([torch.randn(2, 8, 16)], [(4, 4)], loc, torch.rand(1, 2, 2, 1, 3))


def test_level_start_index():
    assert dp.level_start_index([(80, 80), (40, 40), (20, 20)]) == [0, 6400, 8000]
    assert dp.level_start_index([[40, 40], [20, 20]]) == [0, 1600]


def test_image_range_partitions_exactly():
    for n in (0, 1, 7, 64, 255, 256):
        for world in (1, 2, 3, 4, 8):
            ranges = [shard.image_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            for a, b in zip(ranges, ranges[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1
            assert sum(sizes) == n
    with pytest.raises(ValueError):
        shard.image_range(4, 4, 4)


def test_algorithmic_bytes_match_baseline_md():
    # BASELINE.md §3 / SURVEY §8d: S/L shape per image fp32 = 10.95 MB fwd, 20.8 MB bwd
    w = synthetic.WORKLOADS["detrpose_s"]
    f, b = synthetic.algorithmic_bytes(1, w["Lq"], w["H"], w["Dh"], w["shapes"], w["P"], e_v=4, e_o=4)
    assert abs(f / 1e6 - 10.95) < 0.02 and abs(b / 1e6 - 20.8) < 0.05
    f16, _ = synthetic.algorithmic_bytes(1, w["Lq"], w["H"], w["Dh"], w["shapes"], w["P"], e_v=2, e_o=2)
    assert abs(f16 / 1e6 - 6.10) < 0.02
    wn = synthetic.WORKLOADS["detrpose_n"]
    fn, _ = synthetic.algorithmic_bytes(1, wn["Lq"], wn["H"], wn["Dh"], wn["shapes"], wn["P"], e_v=4, e_o=4)
    assert abs(fn / 1e6 - 2.82) < 0.02


def test_synthetic_inputs_are_seeded_and_shaped():
    w = synthetic.WORKLOADS["detrpose_n"]
    a = synthetic.make_inputs(2, 12, w["H"], w["Dh"], w["shapes"], w["P"], seed=3)
    b = synthetic.make_inputs(2, 12, w["H"], w["Dh"], w["shapes"], w["P"], seed=3)
    for k in ("memory", "locations", "attention", "grad_out"):
        assert torch.equal(a[k], b[k])
    assert a["memory"].shape == (2, 2000, 128)
    assert a["locations"].shape == (2, 12, 8, 2, 6, 2)
    assert torch.allclose(a["attention"].sum((-1, -2)), torch.ones(2, 12, 8), atol=1e-5)
    assert a["locations"].min() >= -0.1 and a["locations"].max() <= 1.1
    d = synthetic.make_inputs(1, 4, 8, 16, w["shapes"], 6, seed=1, degenerate=True)
    assert torch.equal(d["locations"][..., 0, :], d["locations"][..., 5, :])


# ---- row f2: value producer patch (host logic only, no kernels) ----
def test_lazy_heads_follows_the_reference_chain_and_falls_back():
    from detrpose_b200 import patch, functional as MF
    mem = torch.randn(2, 10, 64, requires_grad=True)
    sizes = [6, 4]
    lazy = patch._LazyHeads(mem, (8, -1))
    vl = lazy.permute(0, 2, 3, 1).flatten(0, 1).split(sizes, dim=-1)
    assert isinstance(vl, MF.ValueList) and vl.memory is mem and vl.n_heads == 8 and len(vl) == 2
    want = mem.unflatten(2, (8, -1)).permute(0, 2, 3, 1).flatten(0, 1).split(sizes, dim=-1)
    assert all(torch.equal(a, b) for a, b in zip(vl, want))
    assert torch.equal(vl[1], want[1])
    # any other use of the intermediate gets the real tensor
    other = patch._LazyHeads(mem, (8, -1)).permute(0, 1, 3, 2)
    assert isinstance(other, torch.Tensor) and other.shape == (2, 10, 8, 8)
    assert patch._LazyHeads(mem, (8, -1)).shape == (2, 10, 8, 8)
    # gradients flow through the lazily built list like through the reference's expression
    vl[0].sum().backward()
    assert mem.grad is not None and float(mem.grad.abs().sum()) == 2 * 6 * 64


def test_install_value_producer_wraps_and_restores():
    import types
    from detrpose_b200 import patch

    class Transformer:
        def _get_encoder_input(self, feats):
            return feats[0], [[2, 5]], [10]

    mod = types.SimpleNamespace(Transformer=Transformer)
    original = Transformer._get_encoder_input
    patch.install_value_producer(mod)
    patch.install_value_producer(mod)                         # idempotent
    assert Transformer._get_encoder_input is not original
    mem = torch.randn(1, 10, 16)
    out, shapes, sizes = Transformer()._get_encoder_input([mem])
    assert type(out) is torch.Tensor and out is mem           # CPU tensors are left alone
    patch.uninstall_value_producer(mod)
    assert Transformer._get_encoder_input is original


def test_memory_subclass_is_a_plain_tensor_for_everything_but_unflatten():
    from detrpose_b200 import patch
    mem = torch.randn(2, 10, 64, requires_grad=True)
    m = (mem * 1).as_subclass(patch._Memory)
    assert type(m.masked_fill(torch.zeros(2, 10, 1, dtype=torch.bool), 0.0)) is torch.Tensor
    assert type(m + 1) is torch.Tensor and m.size(0) == 2 and m.requires_grad
    v = m.unflatten(2, (8, -1))                               # CPU: the real op
    assert type(v) is torch.Tensor and v.shape == (2, 10, 8, 8)
    v.sum().backward()
    assert mem.grad is not None


def test_kernel_contract_routing_predicate():
    from detrpose_b200 import patch
    loc = torch.zeros(1, 3, 8, 2, 4, 2)
    ok = [torch.zeros(8, 32, 6), torch.zeros(8, 32, 4)]
    assert patch._inside_kernel_contract(ok, loc)
    assert not patch._inside_kernel_contract([torch.zeros(8, 12, 6)], loc)            # Dh not a multiple of 8
    assert not patch._inside_kernel_contract(ok, torch.zeros(1, 3, 8, 2, 20, 2))      # 20 points
    assert not patch._inside_kernel_contract(ok, torch.zeros(1, 3, 8, 9, 4, 2))       # 9 levels
    assert not patch._inside_kernel_contract([v.double() for v in ok], loc)


def test_division_magic_is_exact_in_the_range_the_kernels_use():
    """csrc/msda_common.cuh `div_magic` / `fastdiv`: n // d == (n * ceil(2**32 / d)) >> 32 for every divisor the
    C ABI accepts (L*P <= 8*16, H up to 64, P <= 16) and every numerator the forward forms (< 2**16: thread and
    sample indices of a CTA, head index plus items per CTA)."""
    import numpy as np
    n = np.arange(0, 1 << 16, dtype=np.uint64)
    for d in list(range(2, 130)) + [192, 255, 256, 1000, 4096]:
        magic = np.uint64(((1 << 32) + d - 1) // d)
        assert magic < (1 << 32)
        assert np.array_equal((n * magic) >> np.uint64(32), n // np.uint64(d)), d
    # worst case of the bound n * d < 2**32
    for d in (3, 7, 12, 127):
        top = np.arange((1 << 32) // d - 1000, (1 << 32) // d, dtype=np.uint64)
        magic = np.uint64(((1 << 32) + d - 1) // d)
        assert np.array_equal((top * magic) >> np.uint64(32), top // np.uint64(d)), d
