"""Serving-side product API (vrcoc/session.py): box decode kernel vs the oracle restatement of utils/utils_bbox.py:32-84, and the
pipelined InferenceSession against plain model calls."""
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


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_decode_outputs_vs_oracle(V, dtype):
    from oracle import coc_oracle as O
    g = torch.Generator().manual_seed(3)
    outs = [torch.randn(3, 9, s, s, generator=g).to(dtype) for s in (64, 32, 16)]
    ref = O.decode_outputs([o.double() for o in outs], (512, 512))
    got = V.decode_outputs([o.cuda() for o in outs], (512, 512))
    assert got.shape == (3, 64 * 64 + 32 * 32 + 16 * 16, 9) and got.dtype == torch.float32
    assert rel_err(got, ref) < 1e-5


def test_inference_session_matches_direct_calls(V):
    """two pipeline slots, CUDA graphs, pinned host buffers: results equal the eager model + decode, batch after batch"""
    from test_gpu_parity import _randomised_model
    m = _randomised_model(V, "nano").cuda().to(torch.bfloat16)
    B = 2
    sess = V.InferenceSession(m, batch=B, img=512, slots=2)
    g = torch.Generator().manual_seed(9)
    batches = [(torch.randn(B, 3, 512, 512, generator=g).to(torch.bfloat16).pin_memory(),
                torch.rand(B, 4, 512, 512, generator=g).to(torch.bfloat16).pin_memory()) for _ in range(4)]
    got = []
    for i, (x, r) in enumerate(batches):           # pipelined: collect batch i-1 after submitting batch i
        sess.submit(x, r)
        if i > 0:
            got.append([t.clone() for t in sess.collect()])
    got.append([t.clone() for t in sess.collect()])
    for (x, r), (boxes, cls) in zip(batches, got):
        with torch.no_grad():
            det, seg = m(x.cuda(), r.cuda())
            ref_boxes = V.decode_outputs(det, (512, 512)).cpu()
            ref_cls = seg.argmax(1).to(torch.uint8).cpu()
        assert rel_err(boxes, ref_boxes) < 1e-3          # graph replay vs eager: statistics atomics order only
        assert (cls == ref_cls).float().mean() > 0.999
    with pytest.raises(V.VrcocError):
        sess.collect()


@pytest.mark.parametrize("cuda_graph", [True, False], ids=["graph", "eager"])
def test_concurrent_slots_match_serial_session(V, cuda_graph):
    """concurrent=True: every slot replays on its own compute stream with its own statistics arena (ops.sums_arena.lane), the
    forwards of consecutive batches overlap; results must equal the serial session's, batch after batch, also when the slots are
    hammered back to back without collecting in between (device-resident inputs, readback=False)"""
    from test_gpu_parity import _randomised_model
    m = _randomised_model(V, "nano").cuda().to(torch.bfloat16)
    B = 2
    conc = V.InferenceSession(m, batch=B, img=512, slots=3 if cuda_graph else 2, concurrent=True, cuda_graph=cuda_graph)
    ser = V.InferenceSession(m, batch=B, img=512, slots=2, concurrent=False, cuda_graph=cuda_graph)
    assert conc.concurrent and not ser.concurrent and len({S["lane"] for S in conc.slots}) == len(conc.slots) and len(conc.streams()) == 4
    g = torch.Generator().manual_seed(21)
    batches = [(torch.randn(B, 3, 512, 512, generator=g).to(torch.bfloat16).pin_memory(),
                torch.rand(B, 4, 512, 512, generator=g).to(torch.bfloat16).pin_memory()) for _ in range(6)]

    def drive(sess):
        got = []
        for i, (x, r) in enumerate(batches):
            sess.submit(x, r)
            if i > 0:
                got.append([t.clone() for t in sess.collect()])
        got.append([t.clone() for t in sess.collect()])
        return got

    a, b = drive(conc), drive(ser)
    for (ba, ca), (bb, cb) in zip(a, b):
        assert torch.isfinite(ba).all()
        assert rel_err(ba, bb) < 1e-3                      # statistics atomics order only
        assert (ca == cb).float().mean() > 0.999
    # device-resident, no read-back, no host synchronisation between submits: slot reuse is ordered on the device
    dev_batches = [(x.cuda(), r.cuda()) for x, r in batches]
    for rep in range(3):
        for x, r in dev_batches:
            conc.submit(x, r, readback=False)
    torch.cuda.synchronize()
    last = [t.clone() for t in conc.slots[(conc._next - 1) % len(conc.slots)]["outs"]]
    ser.submit(*batches[-1])
    ref = ser.collect()
    assert rel_err(last[0].float().cpu(), ref[0]) < 1e-3 and (last[1].cpu() == ref[1]).float().mean() > 0.999
