"""-m gpu: the UNMODIFIED reference (staged under oracle/_ref by oracle/stage_ref.py) beside the
patched one on the same B200.

``patch.install()`` rebinds ``Detect.forward`` (effidehead.py:214-301), ``non_max_suppression``
(nms.py:31-130) and ``Inferer.rescale`` (inferer.py:203-228) inside the real package; the flow below
is the post-model flow of ``Inferer.infer`` (inferer.py:82,100).  Stage-wise protocol:
convolutions differ between cuDNN and the CPU by ~1e-6, so every stage is compared on IDENTICAL
inputs -- bit-exact where the arithmetic is fp32 add/sub/mul/div, 1e-5 relative for the sigmoid
columns -- and the whole chain patched-on-GPU vs unpatched-on-CPU within the north-star tolerance.
"""
import numpy as np
import pytest
import torch

from oracle import stage_ref
from _util import assert_rows_equal

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CH = (64, 128, 256)     # LP-s head widths (configs/yololps.py)


@pytest.fixture(scope="module")
def ref():
    if not stage_ref.is_staged():
        pytest.skip("oracle/_ref is not staged (run `python oracle/stage_ref.py` in the build container)")
    torch.backends.cudnn.allow_tf32 = False     # fp32 convolutions: the comparison is against the CPU's
    torch.backends.cuda.matmul.allow_tf32 = False
    return stage_ref.reference()


@pytest.fixture()
def patched(ref):
    from yolo_lp_b200 import patch
    done = patch.install()
    assert {"yolov6.utils.nms.non_max_suppression", "yolov6.core.inferer.Inferer.rescale",
            "yolov6.models.effidehead.Detect.forward"} <= set(done)
    yield ref
    patch.uninstall()


def _feats(B, H, W, seed, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return [torch.rand(B, c, H // s, W // s, generator=g).to(dtype) for c, s in zip(CH, (8, 16, 32))]


def _clone(feats, device=None):
    return [f.clone() if device is None else f.to(device) for f in feats]


def _match_rows(got, want, rtol, atol):
    """Every row of ``got`` has a counterpart in ``want`` (nearest by box) within tolerance."""
    assert got.shape == want.shape, f"kept {got.shape[0]} vs {want.shape[0]}"
    used = set()
    for r in got:
        d = np.abs(want[:, :4] - r[:4]).sum(1)
        j = int(np.argmin(d))
        assert j not in used, "two detections matched the same reference row"
        used.add(j)
        np.testing.assert_allclose(r[:12], want[j, :12], rtol=rtol, atol=atol)          # pixels
        np.testing.assert_allclose(r[12:20], want[j, 12:20], rtol=rtol, atol=1e-6)      # confidences
        assert np.array_equal(r[20:], want[j, 20:]), "class ids differ"


@pytest.mark.parametrize("B,H,W,conf", [(2, 640, 640, 0.3), (1, 384, 640, 0.2), (3, 320, 320, 0.25)])
def test_patched_detect_nms_rescale_against_the_real_reference(ref, B, H, W, conf):
    from yolo_lp_b200 import patch
    head_cpu = ref.build_head(CH, rerandomise=True, seed=3)
    head_gpu = ref.build_head(CH, rerandomise=True, seed=3).to(DEV)
    feats = _feats(B, H, W, seed=H + B)
    nms, rescale = ref.nms.non_max_suppression, ref.inferer.Inferer.rescale
    src_shape = (1160, 720, 3)

    with torch.no_grad():
        # ---- unpatched, everything on the CPU: the reference as the parity target
        cpu_head = head_cpu(_clone(feats))
        cpu_rows = nms(cpu_head.clone(), conf, 0.45, max_det=300)
        # ---- unpatched on the GPU: the reference's own CUDA route, same cuDNN convolutions as ours
        ref_gpu_head = head_gpu(_clone(feats, DEV))
        assert ref.effidehead.Detect.forward.__module__ == "yolov6.models.effidehead"
        patch.install()
        try:
            assert ref.effidehead.Detect.forward.__module__ == "yolo_lp_b200.patch"
            our_head = head_gpu(_clone(feats, DEV))                                  # lp_detect_decode_f32
            our_rows = ref.nms.non_max_suppression(our_head, conf, 0.45, max_det=300)   # lp_nms_f32 (rebound name)
            our_final = []
            for r in our_rows:                                                       # inferer.py:100
                det = r.clone()
                det[:, :12] = ref.inferer.Inferer.rescale((H, W), det[:, :12], src_shape).round()
                our_final.append(det)
        finally:
            patch.uninstall()
        assert ref.nms.non_max_suppression is nms

    # stage 1, Detect.forward on identical conv outputs: geometry columns bit-exact, sigmoids 1e-5
    a, b = our_head.cpu().numpy(), ref_gpu_head.cpu().numpy()
    assert a.shape == b.shape == tuple(cpu_head.shape)
    assert np.array_equal(a[..., :13].view(np.uint32), b[..., :13].view(np.uint32)), "box / obj / corner columns differ"
    np.testing.assert_allclose(a[..., 13:], b[..., 13:], rtol=1e-5, atol=0)
    np.testing.assert_allclose(a, cpu_head.numpy(), rtol=1e-3, atol=1e-2)   # cuDNN vs CPU convolutions (pixels)

    # stage 2, non_max_suppression on the identical head tensor: unpatched CPU reference, bit-exact
    want_rows = nms(our_head.cpu().clone(), conf, 0.45, max_det=300)
    assert sum(int(w.shape[0]) for w in want_rows) >= 10 * B, "test input keeps too few detections to mean anything"
    for i in range(B):
        assert_rows_equal(our_rows[i].cpu().numpy(), want_rows[i].numpy(), f"patched NMS vs reference NMS [{i}]")

    # stage 3, Inferer.rescale + round on identical rows: bit-exact
    for i in range(B):
        want = want_rows[i].clone()
        want[:, :12] = rescale((H, W), want[:, :12], src_shape).round()
        assert_rows_equal(our_final[i].cpu().numpy(), want.numpy(), f"patched rescale vs reference rescale [{i}]")

    # whole chain, patched on the GPU vs unpatched on the CPU from the same features (north star: kept
    # counts equal, values within 1e-5 relative -- plus the convolutions' own 1e-6-level difference)
    for i in range(B):
        _match_rows(our_rows[i].cpu().numpy(), cpu_rows[i].numpy(), rtol=1e-3, atol=1e-2)


def test_true_random_init_lp_s_head_at_640(ref):
    """Config 1's literal wording: YOLO-LP-s random-init head, 640x640, batch 1.  initialize_biases
    zeroes the prediction convs (effidehead.py:94-154), so every class score is sigmoid(-log(99)) and
    every anchor ties: the kept set is decided by NMS order alone (ascending anchor)."""
    from yolo_lp_b200 import patch
    head = ref.build_head(CH, rerandomise=False, seed=1).to(DEV)
    feats = _feats(1, 640, 640, seed=5)
    with torch.no_grad():
        patch.install()
        try:
            our_head = head(_clone(feats, DEV))
            our_rows = ref.nms.non_max_suppression(our_head, 0.001, 0.45, max_det=1000)   # tools/infer.py:26 max_det
        finally:
            patch.uninstall()
        want = ref.nms.non_max_suppression(our_head.cpu().clone(), 0.001, 0.45, max_det=1000)
    assert our_head.shape == (1, 8400, 290)
    assert want[0].shape[0] > 100
    assert_rows_equal(our_rows[0].cpu().numpy(), want[0].numpy(), "degenerate random-init head")


def test_use_dfl_head_stays_with_the_reference(patched):
    """the upstream yolov6m/l heads (use_dfl=True, reg_max=16) must keep working after
    install(): the patched forward hands them back to the original."""
    ref = patched
    e = ref.effidehead
    ch = [0] * 11
    ch[6], ch[8], ch[10] = 32, 64, 128
    torch.manual_seed(0)
    dfl = e.Detect(31, 24, 37, 3, head_layers=e.build_effidehead_layer(ch, 1, 31, 24, 37, reg_max=16, num_layers=3),
                   use_dfl=True, reg_max=16).eval().to(DEV)
    feats = [torch.rand(2, c, 160 // s, 160 // s, device=DEV) for c, s in zip((32, 64, 128), (8, 16, 32))]
    from yolo_lp_b200 import patch
    with torch.no_grad():
        got = dfl(_clone(feats))
        patch.uninstall()
        want = dfl(_clone(feats))
    assert got.shape == (2, 525, 290) and torch.equal(got, want)


def test_half_mode_dtype_flow_and_agreement_with_the_reference_cuda_route(ref):
    """--half (inferer.py:46-50): model.half() on the GPU.  The reference's head tensor is fp32 even
    then (fp32 anchors promote dist2bbox / dist2cor and torch.cat promotes the half sigmoids), so its
    NMS runs in fp32.  The drop-in keeps that flow (lp_detect_decode_half_scores_f32): box / corner
    columns are the same fp32 arithmetic on the same half conv outputs -- bit-identical to the
    reference's CUDA route -- and the class columns are sigmoids rounded to half like the reference's
    (equal except where the 1e-5 sigmoid difference straddles a half rounding boundary)."""
    from yolo_lp_b200 import patch
    B, H, W, conf = 2, 640, 640, 0.3
    head = ref.build_head(CH, rerandomise=True, seed=3).to(DEV).half()
    feats = _feats(B, H, W, seed=9, dtype=torch.float16)
    with torch.no_grad():
        their_head = head(_clone(feats, DEV))
        their_rows = ref.nms.non_max_suppression(their_head.clone(), conf, 0.45, max_det=300)
        patch.install()
        try:
            our_head = head(_clone(feats, DEV))
            our_rows = ref.nms.non_max_suppression(our_head, conf, 0.45, max_det=300)
        finally:
            patch.uninstall()
    assert our_head.dtype == their_head.dtype == torch.float32 and our_head.shape == their_head.shape
    assert all(r.dtype == torch.float32 and r.is_cuda for r in our_rows) and all(r.dtype == torch.float32 for r in their_rows)
    a, b = our_head.cpu().numpy(), their_head.cpu().numpy()
    assert np.array_equal(a[..., :13].view(np.uint32), b[..., :13].view(np.uint32)), "box / obj / corner columns differ"
    assert (a[..., 13:] == b[..., 13:]).mean() >= 0.99
    np.testing.assert_allclose(a[..., 13:], b[..., 13:], rtol=2.0 ** -10, atol=0)  # at most one half ulp
    # the NMS on OUR head tensor: patched GPU vs the reference on the CPU, bit-exact (stage-wise protocol)
    want_rows = ref.nms.non_max_suppression(our_head.cpu().clone(), conf, 0.45, max_det=300)
    for i in range(B):
        assert_rows_equal(our_rows[i].cpu().numpy(), want_rows[i].numpy(), f"half mode NMS [{i}]")
    # nearly all kept detections in common with the reference's own CUDA route
    common = total = 0
    for i in range(B):
        x, y = our_rows[i].cpu().numpy(), their_rows[i].cpu().numpy()
        total += max(x.shape[0], y.shape[0])
        for r in x:
            if y.shape[0] and np.abs(y[:, :4] - r[:4]).max(1).min() == 0.0:
                common += 1
    assert total > 0 and common >= 0.9 * total, f"only {common} of {total} detections in common with the reference's half route"
