#!/usr/bin/env python3
"""Generate the golden fixtures in this directory by RUNNING THE REFERENCE.

Run in the build container only (``/root/reference`` is not on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Imports the unmodified reference callables
  yolov6.utils.nms.non_max_suppression          (yolov6/utils/nms.py:31)
  yolov6.models.effidehead.Detect / build_effidehead_layer (effidehead.py:15,304)
  yolov6.assigners.anchor_generator.generate_anchors       (anchor_generator.py:4)
  yolov6.utils.general.dist2bbox / dist2cor                (general.py:29,51)
  yolov6.core.inferer.Inferer.rescale                      (inferer.py:203)
with torch CPU + torchvision 0.26.0 CPU and stores their outputs.  Large inputs
are NOT stored: they are regenerated from ``yolo_lp_b200.synth`` and pinned by a
SHA-256; small/edge inputs and model-derived tensors are stored in full.

Oracle-driving rules: always clone (the reference mutates its
input, nms.py:76) and call it in small chunks so its 10 s wall-clock
``time_limit`` (nms.py:63,126-128) never fires.
"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
sys.path.insert(0, REPO)

import numpy as np
import torch

from yolov6.utils.nms import non_max_suppression as ref_nms
from yolov6.utils.general import dist2bbox as ref_dist2bbox, dist2cor as ref_dist2cor
from yolov6.assigners.anchor_generator import generate_anchors as ref_generate_anchors
from yolov6.models.effidehead import Detect as RefDetect, build_effidehead_layer
from yolov6.core.inferer import Inferer as RefInferer

from yolo_lp_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
torch.manual_seed(0)


def run_ref_nms(pred, conf, iou, max_det, chunk=1):
    outs = []
    for s in range(0, pred.shape[0], chunk):
        outs += ref_nms(pred[s:s + chunk].clone(), conf, iou, max_det=max_det)
    return [o.numpy().astype(np.float32) for o in outs]


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.1f} KiB")


def pack_rows(rows):
    counts = np.array([r.shape[0] for r in rows], np.int64)
    flat = np.concatenate(rows, 0) if len(rows) else np.zeros((0, 28), np.float32)
    return counts, flat.astype(np.float32)


# ----------------------------------------------------------------- seeded synthetic NMS cases
# name -> (config id, image indices, quant, overrides)
SEEDED = {
    "nms_cfg1": (1, [0], None, {}),
    "nms_cfg2": (2, [0, 1, 2, 31], None, {}),
    "nms_cfg3": (3, [0, 100, 255], None, {}),
    "nms_cfg4_dense": (4, [0, 63], None, {}),
    "nms_cfg5_1280": (5, [0, 31], None, {}),
    "nms_cfg2_ties16": (2, [0, 5], 16, {}),
    "nms_cfg4_ties16": (4, [1], 16, {}),
    "nms_eval_default": (2, [3, 4], None, dict(conf=0.03, iou=0.65)),
    "nms_cfg1_maxdet5": (1, [0], None, dict(max_det=5)),
}


def make_seeded():
    for name, (cid, idxs, quant, over) in SEEDED.items():
        cfg = dict(synth.CONFIGS[cid])
        cfg.update(over)
        rows, shas = [], []
        for i in idxs:
            x = synth.synth_image(cfg["A"], cfg["img"], cfg["n_plates"], cfg["n_pos"], cfg["seed"], i, quant=quant)
            shas.append(synth.sha256_of(x))
            rows += run_ref_nms(x[None], cfg["conf"], cfg["iou"], cfg["max_det"])
        counts, flat = pack_rows(rows)
        save(name, config_id=cid, indices=np.array(idxs), quant=np.array(quant or 0),
             conf=np.array(cfg["conf"]), iou=np.array(cfg["iou"]), max_det=np.array(cfg["max_det"]),
             sha256=np.array(shas), counts=counts, rows=flat)
        print("   counts", counts.tolist())


# ----------------------------------------------------------------- small stored-input edge cases
def edge_inputs():
    g = torch.Generator().manual_seed(1234)
    A = 96
    cases = {}

    def base(A=A, obj=None):
        x = synth.synth_image(A, 128, 6, 24, 77, len(cases))
        if obj is not None:
            x[:, 4] = obj
        return x

    cases["plain"] = (base(), 0.25, 0.45, 300)
    cases["none_pass"] = (base(), 0.99, 0.45, 300)                     # 0 survivors
    x = base(); x[:, 13:] *= 0.01; x[17, 13:] = 0.9
    cases["one_pass"] = (x, 0.5, 0.45, 300)                            # 1 survivor
    cases["all_pass_maxdet7"] = (base(), 0.0, 0.45, 7)                 # k > max_det
    cases["obj_random"] = (base(obj=torch.rand(A, generator=g)), 0.1, 0.5, 300)   # obj != 1 (nms.py:76)
    x = base(); x[:, 2:4] = 0.0
    cases["zero_area"] = (x, 0.0, 0.45, 300)                           # 0/0 IoU -> NaN -> never suppressed
    x = base(); x[:, 0:4] = torch.tensor([50.0, 50.0, 20.0, 20.0])
    cases["identical_boxes"] = (x, 0.0, 0.45, 300)                     # IoU 1 everywhere
    # IoU exactly float32(180/400) = 0.44999998 vs thr 0.45 (double) -> kept
    x = base(); x[:, 13:] = 0.0
    x[0, 0:4] = torch.tensor([10.0, 10.0, 20.0, 20.0]); x[0, 13:] = 0.9
    x[1, 0:4] = torch.tensor([4.5, 10.0, 9.0, 20.0]); x[1, 13:] = 0.8
    cases["iou_threshold_edge"] = (x, 0.5, 0.45, 300)
    x = base(); x[:, 13:] = torch.round(x[:, 13:] * 4) / 4
    cases["ties_quant4"] = (x, 0.0, 0.6, 300)
    x = base(A=33)                                                      # odd A, ragged tile
    cases["odd_A33"] = (x, 0.05, 0.45, 300)
    return cases


def make_edges():
    for name, (x, conf, iou, max_det) in edge_inputs().items():
        rows = run_ref_nms(x[None], conf, iou, max_det)
        counts, flat = pack_rows(rows)
        save("nms_edge_" + name, pred=x.numpy(), conf=np.array(conf), iou=np.array(iou),
             max_det=np.array(max_det), counts=counts, rows=flat)
        print("   counts", counts.tolist())
    # a batch mixing empty and non-empty images: checks per-image independence
    xs = torch.stack([edge_inputs()["plain"][0], edge_inputs()["none_pass"][0] * 0.0, edge_inputs()["obj_random"][0]])
    rows = run_ref_nms(xs, 0.2, 0.45, 300, chunk=3)
    counts, flat = pack_rows(rows)
    save("nms_edge_batch3", pred=xs.numpy(), conf=np.array(0.2), iou=np.array(0.45), max_det=np.array(300),
         counts=counts, rows=flat)
    print("   counts", counts.tolist())


# ----------------------------------------------------------------- model-derived (Detect head) cases
CLS = ("pro", "alp", "ad0", "ad1", "ad2", "ad3", "ad4", "ad5")


def build_head(channels=(64, 128, 256), rerandomise=True):
    ch_list = [0] * 11
    ch_list[6], ch_list[8], ch_list[10] = channels
    layers = build_effidehead_layer(ch_list, 1, 31, 24, 37, reg_max=0, num_layers=3)
    head = RefDetect(31, 24, 37, 3, head_layers=layers, use_dfl=False, reg_max=0)
    head.initialize_biases()
    if rerandomise:   # true random-init is degenerate (all preds zeroed)
        with torch.no_grad():
            for name in ("pro_preds", "alp_preds", "ad0_preds", "ad1_preds", "ad2_preds", "ad3_preds",
                         "ad4_preds", "ad5_preds", "reg_preds", "cor_preds"):
                for conv in getattr(head, name):
                    conv.weight.normal_(0.0, 0.5)
    return head.eval()


def run_head(head, B, H, W, channels=(64, 128, 256)):
    feats = [torch.rand(B, c, H // s, W // s) for c, s in zip(channels, (8, 16, 32))]
    raw = [dict() for _ in range(3)]
    hooks = []
    names = [n + "_preds" for n in CLS] + ["reg_preds", "cor_preds"]
    for n in names:
        for lvl, conv in enumerate(getattr(head, n)):
            hooks.append(conv.register_forward_hook(
                lambda m, i, o, lvl=lvl, n=n: raw[lvl].__setitem__(n[:3], o.detach().clone())))
    with torch.no_grad():
        out = head([f.clone() for f in feats])
    for h in hooks:
        h.remove()
    return raw, out


def make_decode():
    for name, rer, B, H, W in (("decode_rerand_96x160", True, 1, 96, 160), ("decode_degenerate_96x160", False, 2, 96, 160),
                               ("decode_rerand_64x64", True, 2, 64, 64)):
        head = build_head(rerandomise=rer)
        raw, out = run_head(head, B, H, W)
        arrays = {}
        for lvl in range(3):
            for k, v in raw[lvl].items():
                arrays[f"l{lvl}_{k}"] = v.numpy()
        conf = 0.001 if (not rer or H == 64) else 0.02
        rows = run_ref_nms(out, conf, 0.45, 300, chunk=B)
        counts, flat = pack_rows(rows)
        save(name, out=out.numpy(), conf=np.array(conf), iou=np.array(0.45), max_det=np.array(300),
             counts=counts, rows=flat, hw=np.array([H, W]), **arrays)
        print("   A", out.shape[1], "counts", counts.tolist())


def make_decode_half():
    """The reference's model.half() forward (inferer.py:46-50) on the CPU: the head AND its inputs are
    halves, the head tensor comes out fp32 (fp32 anchors promote the geometry, torch.cat promotes the
    half sigmoids).  Stored: the half prediction-conv outputs (hooks), the head tensor, the NMS rows."""
    head = build_head(rerandomise=True).half()
    B, H, W = 2, 96, 160
    feats = [torch.rand(B, c, H // s, W // s).half() for c, s in zip((64, 128, 256), (8, 16, 32))]
    raw = [dict() for _ in range(3)]
    hooks = []
    for n in [n + "_preds" for n in CLS] + ["reg_preds", "cor_preds"]:
        for lvl, conv in enumerate(getattr(head, n)):
            hooks.append(conv.register_forward_hook(
                lambda m, i, o, lvl=lvl, n=n: raw[lvl].__setitem__(n[:3], o.detach().clone())))
    with torch.no_grad():
        out = head([f.clone() for f in feats])
    for h in hooks:
        h.remove()
    assert out.dtype == torch.float32 and all(v.dtype == torch.float16 for lv in raw for v in lv.values())
    arrays = {}
    for lvl in range(3):
        for k, v in raw[lvl].items():
            arrays[f"l{lvl}_{k}"] = v.numpy()
    conf = 0.02
    rows = run_ref_nms(out, conf, 0.45, 300, chunk=B)
    counts, flat = pack_rows(rows)
    save("decode_half_96x160", out=out.numpy(), conf=np.array(conf), iou=np.array(0.45), max_det=np.array(300),
         counts=counts, rows=flat, hw=np.array([H, W]), **arrays)
    print("   A", out.shape[1], "counts", counts.tolist())


# ----------------------------------------------------------------- geometry + rescale KATs
def make_geometry():
    g = torch.Generator().manual_seed(5)
    feats = [torch.zeros(1, 1, h, w) for h, w in ((12, 20), (6, 10), (3, 5))]
    ap, st = ref_generate_anchors(feats, torch.tensor([8, 16, 32]), 5.0, 0.5, device="cpu", is_eval=True, mode="af")
    dist = torch.rand((2, ap.shape[0], 4), generator=g) * 6
    cdist = torch.rand((2, ap.shape[0], 8), generator=g) * 6 - 1
    save("geometry", anchor_points=ap.numpy(), stride_tensor=st.numpy(), dist=dist.numpy(), cdist=cdist.numpy(),
         bbox_xyxy=ref_dist2bbox(dist, ap, "xyxy").numpy(), bbox_xywh=ref_dist2bbox(dist, ap, "xywh").numpy(),
         corners=ref_dist2cor(cdist, ap).numpy(), hw=np.array([(12, 20), (6, 10), (3, 5)]))


def make_rescale():
    g = torch.Generator().manual_seed(6)
    arrays = {}
    # (letterboxed H_in, W_in) , (source H0, W0)
    shapes = [((640, 416), (1160, 720)), ((384, 640), (1080, 1920)), ((640, 640), (640, 640)),
              ((224, 640), (375, 1242)), ((640, 640), (1160, 720)), ((1280, 1280), (2000, 3000))]
    for n, (ori, tgt) in enumerate(shapes):
        k = 37
        v = torch.rand((k, 12), generator=g) * torch.tensor([ori[1], ori[0]] * 6) * 1.2 - 20.0
        v[0, :] = 0.5                                                   # exercises round-half-even after /ratio
        res = RefInferer.rescale(ori, v.clone(), tgt + (3,))
        arrays[f"in{n}"] = v.numpy()
        arrays[f"out{n}"] = res.numpy()
        arrays[f"round{n}"] = res.round().numpy()                       # inferer.py:100
        arrays[f"shape{n}"] = np.array([ori[0], ori[1], tgt[0], tgt[1]])
    save("rescale", n=np.array(len(shapes)), **arrays)


def make_txt():
    """--save-txt records: the reference computes them inline in Inferer.infer (inferer.py:92-93,
    103-120), which needs a model and image files; the expressions below are those lines driven
    with the reference's own Inferer.rescale / Inferer.box_convert on reference NMS output."""
    cfg = synth.CONFIGS[1]
    x = synth.synth_image(cfg["A"], cfg["img"], cfg["n_plates"], cfg["n_pos"], cfg["seed"], 0)
    det = ref_nms(x[None].clone(), 0.25, 0.45, max_det=1000)[0]
    arrays = {}
    for n, (letterbox, src) in enumerate((((640, 640), (1160, 720, 3)), ((640, 640), (1080, 1920, 3)), ((640, 640), (37, 53, 3)))):
        d = det.clone()
        d[:, :12] = RefInferer.rescale(letterbox, d[:, :12], src).round()          # inferer.py:100
        gn = torch.tensor(src)[[1, 0, 1, 0]]                                        # :92
        gn_cor = torch.tensor(src)[[1, 0, 1, 0, 1, 0, 1, 0]]                        # :93
        recs, lines = [], []
        for output in d:
            xyxy = output[:4].tolist()
            corners = output[4:12].tolist()
            conf = float(output[12:19].mean())                                      # :113
            xywh = (RefInferer.box_convert(torch.tensor(xyxy).view(1, 4)) / gn).view(-1).tolist()   # :115
            corners_gn = (torch.tensor(corners) / gn_cor).tolist()                  # :116
            cls = output[20:].tolist()                                              # :117
            line = (*cls, *xywh, *corners_gn)                                       # :118
            lines.append(('%g ' * len(line)).rstrip() % line)                       # :120
            recs.append(list(line) + [conf])
        arrays[f"det{n}"] = d.numpy()
        arrays[f"rec{n}"] = np.array(recs, np.float32)
        arrays[f"lines{n}"] = np.array(lines)
        arrays[f"src{n}"] = np.array(src)
    save("txt_records", n=np.array(3), **arrays)


def make_eval():
    """LP metric of Evaler.eval (evaler.py:153-283).  yolov6.core.evaler needs pycocotools at import
    time only (the stock COCO path); a two-class stub on sys.path is enough to import the module and
    call the unmodified method on a stand-in `self`."""
    import tempfile
    import types
    stub = tempfile.mkdtemp()
    os.makedirs(os.path.join(stub, "pycocotools"))
    open(os.path.join(stub, "pycocotools", "__init__.py"), "w").close()
    open(os.path.join(stub, "pycocotools", "coco.py"), "w").write("class COCO: pass\n")
    open(os.path.join(stub, "pycocotools", "cocoeval.py"), "w").write("class COCOeval: pass\n")
    sys.path.insert(0, stub)
    from yolov6.core.evaler import Evaler as RefEvaler
    fake_self = types.SimpleNamespace(eval_speed=lambda task: None)

    g = torch.Generator().manual_seed(11)
    cfg = synth.CONFIGS[2]
    n_img = 48
    preds, targets = [], []
    for i in range(n_img):
        x = synth.synth_image(cfg["A"], cfg["img"], cfg["n_plates"], cfg["n_pos"], 5, i)
        det = ref_nms(x[None].clone(), 0.25, 0.45, max_det=300)[0]
        if i == 7:
            det = det[:0]                                   # an image without predictions
        m = int(torch.randint(0, 4, (1,), generator=g)) if i != 3 else 0     # image 3 has no targets
        m = max(m, 1) if i in (0, 1) else m
        tgt = torch.zeros((m, 20))
        for k in range(m):
            src = det[int(torch.randint(0, max(1, det.shape[0]), (1,), generator=g))] if det.shape[0] else torch.zeros(28)
            jitter = (torch.rand(4, generator=g) - 0.5) * float(torch.rand(1, generator=g)) * 30.0
            tgt[k, 8:12] = src[:4] + jitter
            tgt[k, 12:20] = src[4:12] + (torch.rand(8, generator=g) - 0.5) * float(torch.rand(1, generator=g)) * 40.0
            tgt[k, :8] = src[20:28]
            if torch.rand(1, generator=g) < 0.3:
                tgt[k, int(torch.randint(0, 8, (1,), generator=g))] += 1.0     # one wrong character
        if i == 5 and m > 0:
            tgt[0, 8:12] = det[2, :4]                       # IoU exactly 1.0: falls into no bin (stale iou_idx)
        preds.append(det)
        targets.append(tgt)
    # two "batches", as Evaler.predict returns them
    pb, tb = [preds[:8], preds[8:40], preds[40:]], [targets[:8], targets[8:40], targets[40:]]
    res = RefEvaler.eval(fake_self, pb, tb, None, "val")
    arrays = {f"pred{i}": preds[i].numpy() for i in range(n_img)}
    arrays.update({f"tgt{i}": targets[i].numpy() for i in range(n_img)})
    save("eval_metric", n=np.array(n_img), split=np.array(8), scalars=np.array(res[:5], np.float64),
         mAP_list=np.array(res[5], np.float64), recall_list=np.array(res[6], np.float64), **arrays)
    print("   ", res[:5])


def make_targets():
    """Label prep of Evaler.predict (evaler.py:119-127), inline in the reference: the statements are
    replayed here with the reference's own xywh2xyxy."""
    from yolov6.utils.nms import xywh2xyxy as ref_xywh2xyxy
    g = torch.Generator().manual_seed(21)
    bs, w, h = 6, 640, 384
    T = 17
    targets = torch.zeros((T, 21))
    targets[:, 0] = torch.tensor([0, 0, 1, 3, 3, 3, 4, 4, 5, 5, 5, 5, 0, 1, 2, 2, 4]).float()   # not sorted by image
    targets[:, 1:9] = torch.randint(0, 37, (T, 8), generator=g).float()
    targets[:, 9:13] = torch.rand((T, 4), generator=g)
    targets[:, 13:21] = torch.rand((T, 8), generator=g)
    src = targets.clone()
    targets[:, 9:13] = ref_xywh2xyxy(targets[:, 9:13])                                   # :120
    batch_targets = [torch.zeros((0, 20))] * bs                                          # :121
    for target in targets:                                                               # :122
        for j in range(9, 21, 2):
            target[j] = target[j] * w
            target[j + 1] = target[j + 1] * h
        batch_targets[int(target[0])] = torch.cat((batch_targets[int(target[0])], target[None, 1:]), dim=0)
    arrays = {f"out{i}": batch_targets[i].numpy() for i in range(bs)}
    save("eval_targets", targets=src.numpy(), w=np.array(w), h=np.array(h), bs=np.array(bs), **arrays)


def make_iou_band():
    """Kept counts of the reference for pairs of boxes within ulps of the IoU threshold (inputs are
    regenerated by tests/_util.iou_band_pairs, numpy only)."""
    sys.path.insert(0, os.path.join(REPO, "tests"))
    from _util import IOU_BAND_THRESHOLDS, iou_band_pairs
    out = {}
    for i, iou in enumerate(IOU_BAND_THRESHOLDS):
        pred = iou_band_pairs(iou)
        rows = run_ref_nms(pred, 0.25, iou, 300, chunk=64)
        out["kept_%d" % i] = np.array([r.shape[0] for r in rows], np.int8)
        print("   thr", iou, "suppressed fraction", float((out["kept_%d" % i] == 1).mean()))
    save("iou_band", thresholds=np.array(IOU_BAND_THRESHOLDS, np.float64), **out)


def make_maxnms():
    """More than max_nms = 30000 candidates at the default (nms.py:62,115-116): A = 33600, every anchor
    passes conf 0, distinct scores (tests/_util.maxnms_input) -- the reference's cut then has one answer."""
    sys.path.insert(0, os.path.join(REPO, "tests"))
    from _util import maxnms_input
    x = maxnms_input()
    arrays = {"sha256": np.array(synth.sha256_of(x))}
    for tag, iou, max_det in (("a", 0.45, 300), ("b", 0.9, 1000)):
        rows = run_ref_nms(x[None], 0.0, iou, max_det)
        counts, flat = pack_rows(rows)
        arrays.update({f"iou_{tag}": np.array(iou), f"max_det_{tag}": np.array(max_det), f"counts_{tag}": counts,
                       f"rows_{tag}": flat})
        print("   ", tag, "counts", counts.tolist(), "lowest kept score", float(flat[:, 12:20].mean(1).min()))
    # the cut itself decides this one: nothing suppressed below IoU 0.999 and max_det above the candidate
    # count, so the reference keeps exactly the 30000 best-scored rows (stored as a digest + the tail)
    import hashlib
    rows = run_ref_nms(x[None], 0.0, 0.999, 33600)[0]
    arrays.update(count_c=np.array(rows.shape[0]), sha256_c=np.array(hashlib.sha256(rows.tobytes()).hexdigest()),
                  tail_c=rows[-16:].copy(), lowest_c=np.array(rows[:, 12:20].mean(1).min()))
    print("    c count", rows.shape[0], "lowest kept score", float(rows[:, 12:20].mean(1).min()))
    save("nms_maxnms_33600", **arrays)


if __name__ == "__main__":
    which = sys.argv[1:] or ["seeded", "edges", "decode", "geometry", "rescale", "txt", "eval", "targets", "iou_band", "maxnms", "decode_half"]
    for w in which:
        globals()["make_" + w]()
