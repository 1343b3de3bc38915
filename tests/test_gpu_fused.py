"""GPU parity: the fused K1+R1 kernel is bit-identical to K1 followed by R1 (and therefore to the oracle / torch-CUDA)."""
import numpy as np
import pytest
import torch

from oracle import maskpath_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,target", [((3, 5, 720, 1280), None), ((2, 4, 1080, 1920), None), ((7, 480, 864), None), ((2, 3, 1280, 736), None),
                                          ((2, 3, 96, 160), (54, 96)), ((5, 64, 64), None), ((3, 200, 320), (37, 50)), ((2, 40, 64), (300, 500)),
                                          ((1, 720, 1280), (100, 100)), ((2, 480, 854), None), ((70, 96, 160), (54, 96)),
                                          ((4, 37, 70), None), ((2, 3, 481, 854), None), ((8, 50, 33), (40, 70)), ((3, 50, 33), None), ((16, 21, 1000), (33, 500))])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_fused_equals_two_kernels(shape, target, dtype):
    import sola_b200 as S
    g = torch.Generator().manual_seed(sum(shape))
    from sola_b200 import synth
    n = int(np.prod(shape[:-2]))
    x = synth.smooth_logits(n, shape[-2], shape[-1], seed=sum(shape), device="cpu", cell=max(8, shape[-2] // 6)).view(shape)
    x.view(-1)[::997] = float("nan")
    x.view(-1)[3::991] = 1.0
    x.view(-1)[5::983] = -1.0
    x = x.to(dtype).cuda()
    p_ref, c_ref = S.binarize_pack_stability(x)
    r_ref, a_ref = S.resize_bilinear_bin(p_ref, target, want_area=True)
    p, c, r, a = S.binarize_pack_resize(x, target_shape=target, want_area=True)
    np.testing.assert_array_equal(p.numpy_u32(), p_ref.numpy_u32())
    np.testing.assert_array_equal(c.cpu().numpy(), c_ref.cpu().numpy())
    np.testing.assert_array_equal(r.numpy_u32(), r_ref.numpy_u32())
    np.testing.assert_array_equal(a.cpu().numpy().reshape(-1), a_ref.cpu().numpy().reshape(-1))
    # and against first principles on one plane
    ref = x.reshape(-1, shape[-2], shape[-1])[0].float().cpu().numpy()
    np.testing.assert_array_equal(p.numpy_u32().reshape(-1, shape[-2], p.Wp)[0], O.pack_bits(ref > 0))
    # planes-free variant (only the resized planes + counts are wanted)
    p2, c2, r2 = S.binarize_pack_resize(x, target_shape=target, want_packed=False)
    np.testing.assert_array_equal(r2.numpy_u32(), r_ref.numpy_u32())
    np.testing.assert_array_equal(c2.cpu().numpy(), c_ref.cpu().numpy())


def test_fused_matches_torch_cuda_reference_chain():
    """End of the chain the reference runs on the GPU: (logits > 0).float() -> interpolate(bilinear) > 0.5."""
    import sola_b200 as S
    from sola_b200 import synth
    x = synth.smooth_logits(4, 720, 1280, seed=9, device="cuda", cell=90)
    _, _, r = S.binarize_pack_resize(x)
    m = (x > 0.0).float()
    exp = (torch.nn.functional.interpolate(m[None], size=(540, 960), mode="bilinear") > 0.5)[0]
    np.testing.assert_array_equal(r.numpy_u32(), O.pack_bits(exp.cpu().numpy()))
