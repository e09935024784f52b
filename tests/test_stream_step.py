"""The block-by-block streaming step as one kernel per direction (csrc/stream.cu) against the
eager modules it replaces: OverlapAdd.forward -> RealtimeSTFT/DGT.forward and RealtimeSTFT/DGT.invert -> OverlapAdd.invert
(oadd.py:70-104, stft.py:248-266, dgt.py:284-302).  Same arithmetic, same order.  The synthesis half is bit-identical to the
eager kernels.  The analysis half agrees to rounding (<= 2e-6 of the peak, both equally far from a float64 transform): ptxas
contracts packed `mul.rn.f32x2` + `add.rn.f32x2` pairs into FFMA2 where its scheduling allows, and it fuses the analysis
window into the first butterfly level of the two kernels differently (identical results with a rectangular window)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    assert torch.cuda.is_available()
    from acids_transforms_b200 import transforms, _lib
    _lib.load()
    return transforms


@pytest.mark.parametrize("n_fft,hop,block", [(512, 128, 1024), (1024, 256, 1024), (1024, 256, 768), (2048, 512, 2048),
                                             (4096, 1024, 4096), (256, 64, 192), (128, 32, 96), (64, 16, 48), (32, 8, 24)])
@pytest.mark.parametrize("kind", ["stft", "dgt"])
def test_stream_step_matches_eager_modules(T, n_fft, hop, block, kind):
    from acids_transforms_b200.streaming import StreamStep
    cls = T.RealtimeSTFT if kind == "stft" else T.RealtimeDGT
    oadd, rt = T.OverlapAdd(n_fft, hop).cuda(), cls(n_fft=n_fft, hop_length=hop).cuda()
    g = torch.Generator(device="cuda").manual_seed(n_fft + hop)
    x = 2 * torch.rand((3, 6 * block), generator=g, device="cuda") - 1
    blocks = x.split(block, -1)
    two = StreamStep(oadd, rt, batch_shape=(3,))
    one = StreamStep(oadd, rt, batch_shape=(3,))
    for i, b in enumerate(blocks):
        Xe = rt(oadd(b))
        ye = oadd.invert(rt.invert(Xe))
        X = two.analysis(b)
        y = two.synthesis(Xe)                      # the eager spectrum in: the synthesis half must reproduce the eager bits
        y1 = one.roundtrip(b)
        assert X.shape == Xe.shape and y.shape == ye.shape == b.shape
        peak = float(Xe.abs().max())
        assert float((X - Xe).abs().max()) <= 2e-6 * peak, "analysis, block %d: %g of %g" % (i, float((X - Xe).abs().max()), peak)
        assert torch.equal(y, ye), "synthesis, block %d: %g" % (i, float((y - ye).abs().max()))
        assert float((y1 - ye).abs().max()) <= 4e-6 * max(1.0, float(ye.abs().max())), "round trip, block %d: %g" % (i, float((y1 - ye).abs().max()))
        assert torch.equal(two.tail, oadd.input_buffer) and torch.equal(two.carry, oadd.output_buffer)


def test_stream_step_graph_replay_and_handover(T):
    """graph=True replays the one-kernel round trip from a CUDA graph (state advances in place inside the graph); push() /
    pull() hand the stream over to / from the eager modules."""
    from acids_transforms_b200.streaming import StreamStep
    oadd, rt = T.OverlapAdd(1024, 256).cuda(), T.RealtimeSTFT(n_fft=1024, hop_length=256).cuda()
    g = torch.Generator(device="cuda").manual_seed(5)
    x = 2 * torch.rand((2, 4, 8 * 1024), generator=g, device="cuda") - 1
    blocks = x.split(1024, -1)
    want = [oadd.invert(rt.invert(rt(oadd(b)))).clone() for b in blocks]
    step = StreamStep(oadd, rt, batch_shape=(2, 4), graph=True, block=1024)
    got = [step.roundtrip(b).clone() for b in blocks[:4]]
    tol = 4e-6 * float(torch.stack(want).abs().max())
    for a, b in zip(got, want[:4]):
        assert float((a - b).abs().max()) <= tol
    # hand over to fresh eager modules mid-stream, and back
    oadd2 = T.OverlapAdd(1024, 256).cuda()
    step.oadd = oadd2
    step.push()
    y4 = oadd2.invert(rt.invert(rt(oadd2(blocks[4]))))
    assert float((y4 - want[4]).abs().max()) <= tol
    step.pull()
    assert float((step.roundtrip(blocks[5]) - want[5]).abs().max()) <= tol
    with pytest.raises(RuntimeError):
        step.roundtrip(x[..., :512])


def test_stream_step_rejects_bad_geometry(T):
    from acids_transforms_b200 import ops
    w = torch.hann_window(1024).cuda()
    tail = torch.zeros((2, 768), device="cuda")
    with pytest.raises(RuntimeError):
        ops.stream_analysis(torch.zeros((2, 1000), device="cuda"), w, 1024, 256, tail)      # not a multiple of hop
    with pytest.raises(RuntimeError):
        ops.stream_analysis(torch.zeros((2, 512), device="cuda"), w, 1024, 256, tail)       # shorter than the carried tail
    with pytest.raises(RuntimeError):
        ops.stream_analysis(torch.zeros((2, 1024), device="cuda"), w, 1024, 256, torch.zeros((2, 100), device="cuda"))
