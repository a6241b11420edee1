"""GPU parity of the YOLOv9-E engine against the fp32 PyTorch oracle (standin/yolov9e.py) on seeded weights.

The engine computes in fp16 with fp32 accumulation (the reference's own CUDA path is fp16 autocast,
ref:util/yolov9.py:110-113); tolerances below are for fp16-vs-fp32 drift through ~100 conv layers and are
stated per check."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from omniparser_b200 import synth  # noqa: E402
from omniparser_b200.detector import B200YOLOv9Detector  # noqa: E402
from oracle import ref_restate as R  # noqa: E402
from standin.yolo_weights import yolo_standin  # noqa: E402

DEV = "cuda:0"
TAP_LAYERS = {"x3": "l3", "x5": "l5", "x7": "l7", "x9": "l9", "x19": "l19", "x22": "l22", "x25": "l25", "x28": "l28",
              "x29": "l29", "x32": "l32", "x35": "l35", "x38": "l38", "x41": "l41", "x2": "l2", "x16": "l16", "x18": "l18"}


@pytest.fixture(scope="module")
def setup():
    m = yolo_standin(0)
    det = B200YOLOv9Detector(state_dict=m.state_dict(), device=DEV)
    return m, det


def _oracle_taps(m, x):
    feats = {}
    hooks = [getattr(m, ln).register_forward_hook(lambda mod, i, o, k=k: feats.__setitem__(k, o)) for k, ln in TAP_LAYERS.items()]
    raw = {}
    hooks += [m.detect.cv2[i].register_forward_hook(lambda mod, i_, o, k=i: raw.__setitem__(("box", k), o)) for i in range(3)]
    hooks += [m.detect.cv3[i].register_forward_hook(lambda mod, i_, o, k=i: raw.__setitem__(("cls", k), o)) for i in range(3)]
    with torch.no_grad():
        outs = m(x)
    for h in hooks:
        h.remove()
    return outs, feats, raw


def test_yolo_layers_and_heads(setup):
    m, det = setup
    img = synth.screenshot(3)
    canvas, scale, pl, pt = R.letterbox_numpy(img, 640)
    x = torch.from_numpy(canvas.astype(np.float32).transpose(2, 0, 1) / 255.0).unsqueeze(0)
    outs, feats, raw = _oracle_taps(m, x)
    io = det._get_io(1, img.shape[0], img.shape[1], 640, 300)
    plan = io["plan"]
    plan.canvas.copy_(torch.from_numpy(canvas).to(DEV).unsqueeze(0))
    plan.run()
    torch.cuda.synchronize()
    worst = 0.0
    for k in TAP_LAYERS:
        ref = feats[k]
        got = plan.taps[k].torch().cpu()
        rel = ((got - ref).abs().max() / ref.abs().max()).item()
        rms = ((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
        print(f"tap {k:4s} max-rel {rel:.4f} rms-rel {rms:.5f}")
        worst = max(worst, rms)
    # fp16 activations through ~100 conv layers: relative RMS drift stays below 2% at every tapped layer
    assert worst < 2e-2
    for i in range(3):
        gb = plan.box_out[i].permute(0, 3, 1, 2).cpu()
        gc = plan.cls_out[i].permute(0, 3, 1, 2).cpu()
        eb = (gb - raw[("box", i)]).abs().max().item()
        ec = (gc - raw[("cls", i)]).abs().max().item()
        print(f"head {i}: box-logit max abs err {eb:.4f} (|ref| max {raw[('box', i)].abs().max():.2f}), cls-logit max abs err {ec:.4f}")
        assert eb < 0.25 and ec < 0.25   # logits of magnitude ~10 (class head gain 3): fp16 drift bound ~1-2 %
    # CUDA-graph replay gives the same bytes as the eager launches
    a = [t.clone() for t in plan.cls_out]
    plan.run()
    torch.cuda.synchronize()
    assert all(torch.equal(u, v) for u, v in zip(a, plan.cls_out))


def test_decode_and_nms_match_oracle_given_same_heads(setup):
    """Feed the ORACLE's fp32 head tensors to the CUDA decode+NMS: kept boxes must equal the reference pipeline."""
    from omniparser_b200 import ops
    m, det = setup
    for seed in (0, 3):
        img = synth.screenshot(seed)
        H, W = img.shape[:2]
        canvas, scale, pl, pt = R.letterbox_numpy(img, 640)
        x = torch.from_numpy(canvas.astype(np.float32).transpose(2, 0, 1) / 255.0).unsqueeze(0)
        _, _, raw = _oracle_taps(m, x)
        with torch.no_grad():
            outs = m(x)
        scores, boxes = R.decode_heads(outs)
        b, s, c = R.filter_candidates(scores[0], boxes[0], 0.05, scale, pl, pt)
        keep, kb, ks = R.nms_and_clamp(b, s, c, 0.1, 300, W, H)
        cls_t = [raw[("cls", i)].permute(0, 2, 3, 1).contiguous().to(DEV) for i in range(3)]
        box_t = [raw[("box", i)].permute(0, 2, 3, 1).contiguous().to(DEV) for i in range(3)]
        hw = [(t.shape[1], t.shape[2]) for t in cls_t]
        cap = 8400
        f = lambda *sh, dt=torch.float32: torch.zeros(*sh, dtype=dt, device=DEV)
        cb, cs, cc, cn = f(1, cap, 4), f(1, cap), f(1, cap, dt=torch.int32), f(1, dt=torch.int32)
        dl, ds = f(1, cap, 4), f(1, cap, 1)
        t1 = lambda v: torch.tensor([v], dtype=torch.float32, device=DEV)
        ops.yolo_decode(cls_t, box_t, hw, 1, 1, 0.05, t1(pl), t1(pt), t1(float(np.float32(scale))), cap, cb, cs, cc, cn, dl, ds)
        kidx, ob, osc, oc = f(1, 300, dt=torch.int32), f(1, 300, 4), f(1, 300), f(1, dt=torch.int32)
        ops.batched_nms(cb, cs, cc, cn, 1, cap, 0.1, 300, t1(W), t1(H), kidx, ob, osc, oc)
        torch.cuda.synchronize()
        # dense decode agrees with the oracle to float rounding (expf / softmax order differ by ulps)
        ltrb_ref = torch.cat([outs[2 * i + 1].permute(0, 2, 3, 1).reshape(1, -1, 4) for i in range(3)], 1)
        assert (dl.cpu() - ltrb_ref).abs().max().item() < 1e-4
        assert (ds.cpu()[0, :, 0] - scores[0, :, 0]).abs().max().item() < 1e-6
        n = int(cn.item())
        k = int(oc.item())
        print(f"seed {seed}: candidates gpu {n} / oracle {len(b)}; kept gpu {k} / oracle {len(kb)}")
        assert n == len(b) and k == len(kb)
        assert np.array_equal(kidx[0, :k].cpu().numpy(), keep.numpy().astype(np.int32))
        assert (ob[0, :k].cpu() - kb).abs().max().item() < 2e-3   # pixels; sub-ulp softmax differences times stride/scale


def test_predict_end_to_end(setup):
    """u8 screenshot -> boxes through the public predict(): (1) the integrated GPU pipeline (letterbox -> network ->
    decode -> NMS) equals the REFERENCE post-processing (oracle.ref_restate, CPU) applied to the head tensors the
    GPU network produced: identical kept count/order, boxes within 2e-3 px; (2) informational agreement with the
    all-fp32 oracle.  Exact box identity against the fp32 oracle is not a meaningful target on a seeded stand-in:
    neighbouring anchors of a random network predict unrelated boxes, so a 1e-2 score perturbation (fp16 vs fp32)
    changes which anchor survives NMS; head accuracy itself is asserted in test_yolo_layers_and_heads."""
    from torchvision.ops import box_iou
    m, det = setup
    for seed in (0, 1):
        img = synth.screenshot(seed)
        H, W = img.shape[:2]
        res = det.predict(img, conf=0.05, iou=0.1)[0].boxes
        io = det._get_io(1, H, W, 640, 300)
        plan = io["plan"]
        canvas, scale, pl, pt = R.letterbox_numpy(img, 640)
        assert np.array_equal(plan.canvas[0].cpu().numpy(), canvas)   # device letterbox == Pillow path
        outs = []
        for i in range(3):
            cl = plan.cls_out[i].permute(0, 3, 1, 2).cpu()
            bl = plan.box_out[i].permute(0, 3, 1, 2).cpu().contiguous()
            b, _, h, w = bl.shape
            ltrb = (bl.view(b, 4, 16, h, w).softmax(2) * torch.arange(16.0).view(1, 1, 16, 1, 1)).sum(2)
            outs += [cl, ltrb]
        scores, boxes = R.decode_heads(outs)
        bb, ss, cc = R.filter_candidates(scores[0], boxes[0], 0.05, scale, pl, pt)
        keep, kb, ks = R.nms_and_clamp(bb, ss, cc, 0.1, 300, W, H)
        gb, gs = res.xyxy.cpu(), res.conf.cpu()
        assert len(gb) == len(kb), (len(gb), len(kb))
        assert (gb - kb).abs().max().item() < 2e-3 and (gs - ks).abs().max().item() < 1e-6
        # informational: agreement with the all-fp32 oracle network
        x = torch.from_numpy(canvas.astype(np.float32).transpose(2, 0, 1) / 255.0).unsqueeze(0)
        with torch.no_grad():
            o32 = m(x)
        s32, b32 = R.decode_heads(o32)
        b_, s_, c_ = R.filter_candidates(s32[0], b32[0], 0.05, scale, pl, pt)
        _, kb32, ks32 = R.nms_and_clamp(b_, s_, c_, 0.1, 300, W, H)
        best = box_iou(kb32, gb).max(1).values
        frac = (best > 0.9).float().mean().item()
        print(f"seed {seed}: gpu {len(gb)} boxes, fp32 oracle {len(kb32)} boxes, oracle boxes matched at IoU>0.9: {frac:.2f}")
        assert frac > 0.5


def test_fullres_detect_config1(setup):
    """BASELINE configs[1]: detect only, batch 1, 1920x1080 at full resolution (`scale_img=True, imgsz=(h, w)` ->
    canvas 1088x1920, scale 1.0, no resample; ref:util/utils.py:391-397, ref:util/yolov9.py:52-61).  Head tensors vs the
    fp32 oracle, and predict() vs the reference post-processing of its own heads."""
    m, det = setup
    img = synth.screenshot(5)
    H, W = img.shape[:2]
    imgsz = (H, W)
    res = det.predict(img, conf=0.05, iou=0.1, imgsz=imgsz)[0].boxes
    io = det._get_io(1, H, W, imgsz, 300)
    plan = io["plan"]
    canvas, scale, pl, pt = R.letterbox_numpy(img, imgsz)
    assert canvas.shape == (1088, 1920, 3) and scale == 1.0 and (pl, pt) == (0, 4)
    assert np.array_equal(plan.canvas[0].cpu().numpy(), canvas)
    x = torch.from_numpy(canvas.astype(np.float32).transpose(2, 0, 1) / 255.0).unsqueeze(0)
    _, _, raw = _oracle_taps(m, x)
    outs = []
    for i in range(3):
        gc = plan.cls_out[i].permute(0, 3, 1, 2).cpu()
        gb = plan.box_out[i].permute(0, 3, 1, 2).cpu().contiguous()
        eb = (gb - raw[("box", i)]).abs().max().item()
        ec = (gc - raw[("cls", i)]).abs().max().item()
        print(f"full-res head {i} {tuple(gc.shape)}: box-logit err {eb:.4f}, cls-logit err {ec:.4f}")
        assert eb < 0.3 and ec < 0.3
        b, _, h, w = gb.shape
        outs += [gc, (gb.view(b, 4, 16, h, w).softmax(2) * torch.arange(16.0).view(1, 1, 16, 1, 1)).sum(2)]
    scores, boxes = R.decode_heads(outs)
    bb, ss, cc = R.filter_candidates(scores[0], boxes[0], 0.05, scale, pl, pt)
    keep, kb, ks = R.nms_and_clamp(bb, ss, cc, 0.1, 300, W, H)
    gbx, gs = res.xyxy.cpu(), res.conf.cpu()
    print(f"full-res: {len(bb)} candidates, {len(kb)} kept (gpu {len(gbx)})")
    assert len(gbx) == len(kb)
    assert (gbx - kb).abs().max().item() < 2e-3 and (gs - ks).abs().max().item() < 1e-6


# ------------------------------------------------------------------------------------------------------------------
# Parity-grade detector mode (precision="fp16x3": hi/lo feature maps, three tensor-core products per term) against
# (a) the fp32 oracle network's head tensors and (b) the goldens written by the UNMODIFIED reference
# (oracle/make_golden.py: util/yolov9.py::YOLOv9Detector.predict on the TorchScript stand-in, fp32 on the CPU).
import json  # noqa: E402
from pathlib import Path  # noqa: E402

GOLD = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def setup_x3():
    m = yolo_standin(0)
    det = B200YOLOv9Detector(state_dict=m.state_dict(), device=DEV, precision="fp16x3")
    return m, det


def test_parity_mode_heads_vs_fp32_oracle(setup_x3):
    """Every tapped feature map and all six head tensors of the fp16x3 engine against the fp32 PyTorch oracle on the
    same letterboxed canvas.  Two fp32-grade evaluations of a ~100-layer network differ by their own rounding noise
    amplified through the depth (measured: 1e-5 relative after the first ELAN block, 1.3e-4 at the SPP block); the
    bounds below are 4x the measured drift, 50-100x tighter than the fp16 mode's (2e-2 / 0.25)."""
    m, det = setup_x3
    img = synth.screenshot(3)
    canvas, scale, pl, pt = R.letterbox_numpy(img, 640)
    x = torch.from_numpy(canvas.astype(np.float32).transpose(2, 0, 1) / 255.0).unsqueeze(0)
    outs, feats, raw = _oracle_taps(m, x)
    io = det._get_io(1, img.shape[0], img.shape[1], 640, 300)
    plan = io["plan"]
    assert plan.x3
    plan.canvas.copy_(torch.from_numpy(canvas).to(DEV).unsqueeze(0))
    plan.run()
    torch.cuda.synchronize()
    for k in TAP_LAYERS:
        ref = feats[k]
        got = plan.taps[k].torch().cpu()
        rel = ((got - ref).abs().max() / ref.abs().max()).item()
        print(f"x3 tap {k:4s} max-rel {rel:.2e}")
        assert rel < 5e-4
    for i in range(3):
        gb = plan.box_out[i].permute(0, 3, 1, 2).cpu()
        gc = plan.cls_out[i].permute(0, 3, 1, 2).cpu()
        eb = (gb - raw[("box", i)]).abs().max().item()
        ec = (gc - raw[("cls", i)]).abs().max().item()
        print(f"x3 head {i}: box-logit max abs err {eb:.2e}, cls-logit max abs err {ec:.2e}")
        assert eb < 4e-3 and ec < 4e-3   # logits of magnitude ~10


@pytest.mark.parametrize("name", ["synth_seed0", "synth_seed3_odd", "synth_seed5_3240x2160"])
def test_parity_mode_predict_reproduces_reference_golden(setup_x3, name):
    """predict() of the parity-grade detector returns the boxes the unmodified reference returned: identical kept
    count and ORDER (= identical kept indices after NMS, tie-break rule in DESIGN.md §2), coordinates within 0.5 px
    (measured ~1e-3), scores within 1e-3."""
    _, det = setup_x3
    g = json.loads((GOLD / f"{name}.json").read_text())
    w, h = g["case"]["size"]
    img = synth.screenshot(g["case"]["seed"], w, h)
    res = det.predict(img, conf=g["box_threshold"], iou=0.1)[0].boxes
    gb, gs = res.xyxy.cpu(), res.conf.cpu()
    rb, rs = torch.tensor(g["det_xyxy"], dtype=torch.float32).reshape(-1, 4), torch.tensor(g["det_conf"], dtype=torch.float32)
    from parity_util import match_scored_boxes
    events = match_scored_boxes(gb, gs, rb, rs, px_tol=0.5, score_tol=1e-3)
    perm = list(range(len(rb)))
    for i, j in events:
        perm[i] = j
    db, ds = (gb[perm] - rb).abs().max().item(), (gs[perm] - rs).abs().max().item()
    print(f"{name}: {len(gb)} boxes, max |dxy| {db:.2e} px, max |dconf| {ds:.2e}, tie-class order events {events}")
    assert db <= 0.5 and ds <= 1e-3


@pytest.mark.parametrize("name", ["synth_seed0", "synth_seed3_odd", "synth_seed5_3240x2160"])
def test_fast_mode_flips_vs_reference_golden_are_counted(setup, name):
    """The fast fp16 detector (the reference's own CUDA precision) against the same goldens: box flips caused by fp16
    rounding of near-tied scores are COUNTED and reported (tie-class events), bounded, never hidden."""
    from torchvision.ops import box_iou
    _, det = setup
    g = json.loads((GOLD / f"{name}.json").read_text())
    w, h = g["case"]["size"]
    img = synth.screenshot(g["case"]["seed"], w, h)
    res = det.predict(img, conf=g["box_threshold"], iou=0.1)[0].boxes
    gb = res.xyxy.cpu()
    rb = torch.tensor(g["det_xyxy"], dtype=torch.float32).reshape(-1, 4)
    # boxes in the letterbox padding clamp to zero height (ref:util/yolov9.py:134-135): IoU is 0/0 there, so a counterpart
    # within 2 px on every coordinate counts as well
    iou = torch.nan_to_num(box_iou(rb, gb)) if len(gb) else torch.zeros((len(rb), 0))
    near = (rb[:, None, :] - gb[None, :, :]).abs().amax(-1) <= 2.0 if len(gb) else torch.zeros((len(rb), 0), dtype=torch.bool)
    matched = int(((iou > 0.9) | near).any(1).sum()) if len(gb) else 0
    flips = len(rb) - matched
    print(f"{name}: fp16 detector {len(gb)} boxes vs reference {len(rb)}; matched at IoU>0.9: {matched}; tie-class flips: {flips}")
    assert matched >= 0.5 * len(rb)
