"""GPU: the eval input pipeline kernel (light_head_preprocess_for_eval/_for_test, WARP_RESIZE) against the numpy
restatement -- fp32 with one rounding per operation on both sides, so the bar is bit-exact."""
import numpy as np
import pytest
import torch

from oracle import preprocess as op

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,out", [((375, 500), (480, 480)), ((500, 333), (480, 480)), ((97, 131), (160, 160)),
                                       ((480, 480), (480, 480)), ((1200, 1600), (800, 800)), ((31, 17), (64, 48))])
def test_preprocess_matches_oracle(shape, out):
    import xdet_b200  # noqa: F401
    from xdet_b200.preprocessing import common_preprocessing as cp
    rng = np.random.default_rng(shape[0] * 7 + shape[1])
    img = rng.integers(0, 256, (shape[0], shape[1], 3), dtype=np.uint8)
    ref = op.preprocess_for_test(img, out)
    got = cp.light_head_preprocess_for_test(torch.from_numpy(img).cuda(), out)
    torch.cuda.synchronize()
    assert got.shape == (3,) + out
    assert np.array_equal(got.cpu().numpy().view(np.int32), ref.view(np.int32))
    # value range of whitened pixels: [0 - mean, 2 - mean]
    assert float(got.min()) >= -1.0 and float(got.max()) <= 1.2


def test_eval_wrapper_and_batch_slice():
    import xdet_b200  # noqa: F401
    from xdet_b200.preprocessing import common_preprocessing as cp
    rng = np.random.default_rng(3)
    batch = torch.zeros((2, 3, 160, 160), device="cuda")
    imgs = [rng.integers(0, 256, (120, 90, 3), dtype=np.uint8), rng.integers(0, 256, (64, 200, 3), dtype=np.uint8)]
    labels = torch.tensor([3, 7, 9], device="cuda")
    bboxes = torch.rand((3, 4), device="cuda")
    diff = torch.tensor([0, 1, 0], device="cuda")
    for i, im in enumerate(imgs):
        img, lab, bb, bbox_img = cp.light_head_preprocess_for_eval(torch.from_numpy(im).cuda(), labels, bboxes, (160, 160),
                                                                   difficults=diff, out=batch[i])
        assert img.data_ptr() == batch[i].data_ptr() and lab.tolist() == [3, 9] and bb.shape == (2, 4)
        assert bbox_img.tolist() == [0.0, 0.0, 1.0, 1.0]
        assert np.array_equal(batch[i].cpu().numpy().view(np.int32), op.preprocess_for_test(im, (160, 160)).view(np.int32))
    with pytest.raises(ValueError):
        cp.light_head_preprocess_for_test(torch.zeros((4, 4), dtype=torch.uint8, device="cuda"), (8, 8))
