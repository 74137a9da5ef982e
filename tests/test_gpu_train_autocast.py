"""One training step of the whole model under torch.autocast, the way the reference's fit loop runs it (utils/utils_fit.py:86-116:
`with autocast(): outputs = model_train(images, radars)`, fp16 by default, train.py:56): fp32 master weights, the native autograd
Functions computing the autocast regions in bf16 (ops.amp_function), gradients back in fp32 on every parameter."""
import pytest
import torch

from golden_util import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def V():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import vrcoc
    return vrcoc


@pytest.mark.parametrize("amp_dtype", [torch.float16, torch.bfloat16], ids=["fp16_autocast", "bf16_autocast"])
def test_training_step_under_autocast(V, amp_dtype):
    from test_gpu_parity import _randomised_model
    m = _randomised_model(V, "nano").cuda().train()                      # fp32 master weights, BatchNorm in batch-statistics mode
    g = torch.Generator().manual_seed(17)
    x = torch.randn(2, 3, 512, 512, generator=g).cuda()
    r = torch.rand(2, 4, 512, 512, generator=g).cuda()
    opt = torch.optim.SGD(m.parameters(), lr=1e-3, momentum=0.9)
    before = {n: p.detach().clone() for n, p in m.named_parameters() if p.numel()}
    with torch.autocast("cuda", dtype=amp_dtype):
        det, seg = m(x, r)
        loss = sum(o.float().square().mean() for o in det) + seg.float().square().mean()
    assert all(torch.isfinite(o.float()).all() for o in det) and torch.isfinite(seg.float()).all() and torch.isfinite(loss)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    n_grad = 0
    for n, p in m.named_parameters():
        if p.numel() == 0:
            continue
        assert p.grad is not None, n
        assert p.grad.dtype == torch.float32 and torch.isfinite(p.grad).all(), n
        n_grad += int(p.grad.abs().sum() > 0)
    assert n_grad > 0.9 * len(before)                                    # the loss reaches (nearly) every parameter
    opt.step()
    moved = sum(int(not torch.equal(before[n], p.detach())) for n, p in m.named_parameters() if p.numel())
    assert moved > 0.9 * len(before)
    # the autocast forward agrees with the plain fp32 forward of the same (pre-step) weights to bf16 accuracy
    m2 = _randomised_model(V, "nano").cuda().train()
    with torch.no_grad():
        det32, seg32 = m2(x, r)
    assert rel_err(seg.float(), seg32.float()) < 5e-2
    for a, b in zip(det, det32):
        assert rel_err(a.float(), b.float()) < 5e-2
