"""CPU tests of the host side: the C-ABI library loads and exports exactly what include/ssdhot.h
declares, argument validation happens before any launch, priors / packing / sharding logic."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

import _util as U
from oracle import ssd_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ssdhot.h")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"SSDHOT_API\s+[\w\s\*]+?\b(ssdhot_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import ssdhot
    from ssdhot import _lib
    names = declared_symbols()
    assert len(names) >= 15
    handle = ssdhot.lib()
    for n in names:
        assert hasattr(handle, n), f"{n} declared in ssdhot.h but not exported"
        assert n in _lib.PROTOTYPES, f"{n} has no ctypes prototype"
    assert sorted(_lib.PROTOTYPES) == names
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l and "ssdhot_" in l)
    assert exported == names, "exported symbols differ from the header"
    assert handle.ssdhot_abi_version() == 2
    assert handle.ssdhot_status_string(-2).decode().startswith("a size")


def test_library_is_sm100a_only():
    from ssdhot import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_validation_happens_before_launch_no_gpu_needed():
    import ssdhot
    L = ssdhot.lib()
    assert L.ssdhot_decode(None, None, 8, 0.1, 0.2, None, None) == -1
    assert L.ssdhot_prior_tables(16, 0, 16, 16, None) == -2
    assert L.ssdhot_prior_tables(16, 8732, 20, 16, None) == -5
    assert L.ssdhot_predict(16, 8732, 16, 16, 1, 6, 1.0, 0.45, 200, 0, 0, 0.1, 0.2, 300.0, 300.0,
                            16, 16, 16, None, 16, 16, None) == -3          # score_thresh must be < 1
    assert L.ssdhot_predict(16, 8732, 16, 16, 1, 6, 0.1, 0.0, 200, 0, 0, 0.1, 0.2, 300.0, 300.0,
                            16, 16, 16, None, 16, 16, None) == -3          # nms_thresh must be > 0
    assert L.ssdhot_predict(16, 8732, 16, 16, 1, 1, 0.1, 0.5, 200, 0, 0, 0.1, 0.2, 300.0, 300.0,
                            16, 16, 16, None, 16, 16, None) == -2          # C >= 2
    assert L.ssdhot_loss_workspace_bytes(256, 8732, 20) > 0 and L.ssdhot_match_workspace_bytes(256, 20) > 0 and L.ssdhot_predict_workspace_bytes(256, 8732, 6) > 0
    # head-direct entry points: six 16-byte-aligned pointers per branch, a known layout, C == 6
    six = (ctypes.c_void_p * 6)(*([16] * 6))
    five = (ctypes.c_void_p * 6)(*([16] * 5 + [None]))
    odd = (ctypes.c_void_p * 6)(*([16] * 5 + [20]))
    ptr = lambda a: ctypes.cast(a, ctypes.c_void_p)
    heads = lambda loc, conf, layout, C: L.ssdhot_predict_heads(16, ptr(loc), ptr(conf), layout, 1, C, 0.1, 0.45, 200, 0, 0, 0.1, 0.2,
                                                                  300.0, 300.0, 16, 16, 16, None, 16, 16, 3, None, None)
    assert heads(six, six, 7, 6) == -3                                      # unknown layout
    assert heads(six, six, 0, 21) == -2                                     # other class counts: pack first
    assert heads(six, five, 0, 6) == -1 and heads(six, odd, 1, 6) == -5     # missing / misaligned head
    assert L.ssdhot_predict_heads(16, None, ptr(six), 0, 1, 6, 0.1, 0.45, 200, 0, 0, 0.1, 0.2, 300.0, 300.0,
                                  16, 16, 16, None, 16, 16, 3, None, None) == -1
    assert L.ssdhot_multibox_loss_heads_fwd(16, 16, 16, 1, 16, 16, 16, 1, 4, 300.0, 300.0, ptr(six), ptr(six), 0, 21,
                                            0.5, 0.1, 0.2, 3.0, 16, 16, None, None, None, None, None, None) == -2
    assert L.ssdhot_multibox_loss_heads_fwd(16, 16, 16, 0, 16, 16, 16, 1, 4, 300.0, 300.0, ptr(six), ptr(six), 0, 6,
                                            0.5, 0.1, 0.2, 3.0, 16, 16, None, None, None, None, None, None) == -2    # needs the SSD300 layout
    assert L.ssdhot_multibox_loss_heads_bwd(16, 16, 16, 1, 300.0, 300.0, ptr(six), ptr(six), 5, 6, 0.1, 0.2, 16, 16, 16,
                                            ptr(six), ptr(six), None) == -3
    # the backward's class record sel_cls is an int8: more than 127 classes are refused when it is requested (forward-only is fine up to 256)
    fwd = lambda C, sel: L.ssdhot_multibox_loss_fwd(16, 16, 16, 8732, 0, 16, 16, 16, 1, 4, 300.0, 300.0, 16, 16, C,
                                                     0.5, 0.1, 0.2, 3.0, 16, 20, sel, None, None, None, None, None)
    assert fwd(200, 16) == -2 and fwd(200, None) == -5 and fwd(127, 16) == -5      # (-5: past the shape checks, stopped by the misaligned workspace)
    # key hand-off buffer: B flags (256-byte padded) + B x ceil4(P / 2) words
    assert L.ssdhot_share_bytes(256, 8732) == 1024 + 256 * 4368 * 4 and L.ssdhot_share_bytes(0, 8732) == 0
    assert L.ssdhot_share_reset(None, 4, None) == -1 and L.ssdhot_share_reset(16, 0, None) == -2
    assert L.ssdhot_mined_ce_fwd(16, 16, 16, 1, 8732, 128, 3.0, 16, 16, 16, None) == -2
    assert L.ssdhot_multibox_loss_bwd(16, 8732, 16, 16, 1, 300.0, 300.0, 16, 16, 128, 0.1, 0.2, 16, 16, 16, 16, 16, None) == -2
    # peer all-reduce: rank / world / lag are checked before anything is launched
    one = (ctypes.c_void_p * 1)(16)
    assert L.ssdhot_allreduce_sums_peer(16, ptr(one), 0, 9, 0, None, None) == -2
    assert L.ssdhot_allreduce_sums_peer(16, ptr(one), 1, 1, 0, None, None) == -2
    assert L.ssdhot_allreduce_sums_peer(16, ptr(one), 0, 1, 2, None, None) == -3
    assert L.ssdhot_allreduce_sums_peer(None, ptr(one), 0, 1, 0, None, None) == -1
    assert L.ssdhot_peer_mailbox_bytes() >= 4 * 8 * 4 * 8 + 8
    assert ssdhot.launch_count() == 0


def test_no_cpu_fallback():
    import ssdhot
    with pytest.raises(ssdhot.SsdhotError):
        ssdhot.decode_ssd(torch.zeros((4, 4)), torch.zeros((4, 4)), (0.1, 0.2))
    with pytest.raises(ssdhot.SsdhotError):
        ssdhot.PriorSet(torch.zeros((8732, 4)))
    with pytest.raises(ssdhot.SsdhotError):
        ssdhot.iou_nms(torch.rand((4, 4)), torch.rand((4,)), 0.5)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "automotive-ssd-object-detection_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(base, f)).read()
                assert "oracle" not in txt.lower() or f == "synth.py", f"{f} mentions the oracle"


def test_default_boxes_match_reference_hash():
    import ssdhot
    pri = ssdhot.default_boxes()
    assert pri.shape == (8732, 4) and U.sha(pri) == O.PRIORS_SHA256
    assert torch.equal(pri, O.default_boxes())


def test_pack_targets_ragged_and_empty():
    import ssdhot
    t = [{"boxes": torch.rand((3, 4)), "labels": torch.tensor([0, 1, 2])},
         {"boxes": torch.zeros((0, 4)), "labels": torch.zeros((0,), dtype=torch.int64)},
         {"boxes": torch.rand((1, 4)), "labels": torch.tensor([4])}]
    p = ssdhot.pack_targets(t, "cpu")
    assert p.offsets.tolist() == [0, 3, 3, 4] and p.max_gt == 3 and p.n_img == 3
    assert p.boxes.shape == (4, 4) and p.labels.dtype == torch.int64 and p.offsets.dtype == torch.int32
    assert torch.equal(p.boxes[3], t[2]["boxes"][0])
    e = ssdhot.pack_targets(t[1:2], "cpu")
    assert e.offsets.tolist() == [0, 0] and e.max_gt == 0
    from ssdhot import synth
    b, l, o = synth.pack_targets(t)
    assert torch.equal(b, p.boxes) and torch.equal(l, p.labels) and torch.equal(o, p.offsets)


def test_collate_detection_packs_one_buffer():
    """SSD_trainer.collate_detection drop-in (TR:806-813): images stacked, the ragged ground truth in ONE byte buffer
    [offsets | boxes | labels] whose views are what the kernels read; as_list() gives the reference's List[Dict] back."""
    import ssdhot
    g = torch.Generator().manual_seed(5)
    counts = [3, 0, 1, 7, 0]
    batch = []
    for c in counts:
        tgt = {"boxes": torch.rand((c, 4), generator=g) * 300, "labels": torch.randint(0, 5, (c,), generator=g),
               "image_id": torch.tensor([len(batch)])}
        batch.append((torch.rand((3, 8, 8), generator=g), tgt))
    images, packed = ssdhot.collate_detection(batch)
    assert images.shape == (5, 3, 8, 8) and torch.equal(images[3], batch[3][0])
    assert isinstance(packed, ssdhot.PackedTargets) and packed.n_img == 5 and packed.max_gt == 7 and packed.counts == counts
    assert packed.offsets.tolist() == [0, 3, 3, 4, 11, 11] and packed.offsets.dtype == torch.int32
    box_off, lab_off, nbytes = ssdhot.PackedTargets.layout(5, 11)
    assert packed.buffer.numel() == nbytes and box_off % 16 == 0 and lab_off % 8 == 0
    base = packed.buffer.data_ptr()
    assert packed.offsets.data_ptr() == base and packed.boxes.data_ptr() == base + box_off and packed.labels.data_ptr() == base + lab_off
    back = packed.as_list()
    for (_, tgt), got in zip(batch, back):
        assert torch.equal(got["boxes"], tgt["boxes"].reshape(-1, 4)) and torch.equal(got["labels"], tgt["labels"])
    # the same content as pack_targets of the reference's collate output, and idempotent
    ref = ssdhot.pack_targets([t for _, t in batch], "cpu")
    assert torch.equal(ref.boxes, packed.boxes) and torch.equal(ref.labels, packed.labels) and torch.equal(ref.offsets, packed.offsets)
    assert ssdhot.pack_targets(packed, "cpu") is packed
    # a batch without a single box
    _, empty = ssdhot.collate_detection([(torch.zeros((3, 4, 4)), {"boxes": torch.zeros((0, 4)), "labels": torch.zeros((0,), dtype=torch.int64)})] * 2)
    assert empty.offsets.tolist() == [0, 0, 0] and empty.max_gt == 0 and empty.total == 0 and len(empty.as_list()) == 2
    assert empty.as_list()[0]["boxes"].shape == (0, 4)
    # through a DataLoader (worker processes collate, the loader's pin thread would pin): same batches
    ds = [b for b in batch]
    loader = torch.utils.data.DataLoader(ds, batch_size=2, collate_fn=ssdhot.collate_detection, num_workers=0)
    got = [(im.shape[0], pk.offsets.tolist()) for im, pk in loader]
    assert got == [(2, [0, 3, 3]), (2, [0, 1, 8]), (1, [0, 0])]


def test_synth_is_deterministic_and_shaped():
    from ssdhot import synth
    a, b = synth.config(2, batch=3), synth.config(2, batch=3)
    assert torch.equal(a["loc_all"], b["loc_all"]) and torch.equal(a["conf_infer"], b["conf_infer"])
    assert a["loc_all"].shape == (3, 8732, 4) and a["conf_train"].shape == (3, 8732, 6)
    for t in a["targets"]:
        bx = t["boxes"]
        assert 1 <= bx.shape[0] <= 20 and bool((bx[:, 2:] - bx[:, :2] >= 1.0 - 1e-4).all())
        assert bool((bx >= 0).all()) and bool((bx <= 300).all())
    c = synth.config(3, batch=1, dedup=True)["conf_infer"]
    s = c.softmax(-1)[0, :, 1:].reshape(-1)
    assert s.unique().numel() == s.numel()
    r0, r1 = synth.config(4, batch=2, seed_offset=0), synth.config(4, batch=2, seed_offset=1)
    assert not torch.equal(r0["loc_all"], r1["loc_all"])


def test_shard_ranges_cover_the_batch():
    from ssdhot import dist as D
    for n, w in ((4096, 8), (10, 3), (2, 4), (256, 1)):
        spans = [D.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


def _gloo_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, os.path.join(ROOT, "automotive-ssd-object-detection_b200"))
    sys.path.insert(0, ROOT)
    from ssdhot import dist as D, synth
    from oracle import ssd_oracle as O
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = synth.config(2, batch=6)
    pri, pri_xyxy = O.prior_tables()
    lo, hi = D.shard_range(6, rank, world)
    tg = cfg["targets"][lo:hi]
    # per-rank un-normalised sums of the shard, as ssdhot_multibox_loss_fwd writes them
    pos, loc_pm, cls_t = O.batch_targets(pri, pri_xyxy, tg, 300, 300, 0.5)
    n_img = pos.sum(1)
    one = torch.ones(())
    s_loc = torch.nn.functional.smooth_l1_loss(cfg["loc_all"][lo:hi][pos], loc_pm, reduction="sum")
    s_ce = O.mined_ce_loss(cfg["conf_train"][lo:hi], cls_t, pos, n_img, one, 3.0)
    sums = torch.tensor([s_loc.item(), s_ce.item(), float(n_img.sum())], dtype=torch.float64)
    D.combine_sums(sums)
    l_loc, l_conf = D.losses_from_sums(sums)
    out[rank] = (l_loc.item(), l_conf.item())
    dist.destroy_process_group()


def test_sharded_loss_equals_single_process_gloo_world2():
    """world_size-2 gloo run of the sharded path's only exchange: all-reduced partial sums give the
    same losses as the un-sharded reference computation on the concatenated batch (1e-5)."""
    import torch.multiprocessing as mp
    from ssdhot import synth
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_gloo_worker, args=(2, port, out), nprocs=2, join=True)
    cfg = synth.config(2, batch=6)
    pri, pri_xyxy = O.prior_tables()
    want_loc, want_conf = O.train_half(pri, pri_xyxy, cfg["loc_all"], cfg["conf_train"], cfg["targets"], 0.5, 3.0)
    assert out[0] == out[1]
    assert abs(out[0][0] - want_loc.item()) <= 1e-5 * abs(want_loc.item())
    assert abs(out[0][1] - want_conf.item()) <= 1e-5 * abs(want_conf.item())


def test_ssd300_layout_check_host():
    """ssdhot_ssd300_layout_host is a pure host function: true for the reference's default boxes,
    false for anything that breaks the grid structure the box-centric matching kernel relies on."""
    import ssdhot
    L = ssdhot.lib()
    pri = ssdhot.default_boxes().contiguous()
    assert L.ssdhot_ssd300_layout_host(pri.data_ptr(), 8732) == 1
    assert L.ssdhot_ssd300_layout_host(pri.data_ptr(), 8731) == 0
    assert L.ssdhot_ssd300_layout_host(None, 8732) == 0
    for idx, col, delta in ((5000, 0, 1e-3), (17, 2, 1e-4), (8731, 3, 0.5), (0, 1, -2e-5)):
        bad = pri.clone()
        bad[idx, col] += delta
        assert L.ssdhot_ssd300_layout_host(bad.data_ptr(), 8732) == 0, (idx, col)
    perm = pri.clone()
    perm[[0, 1]] = perm[[1, 0]]            # two shapes of one cell swapped
    assert L.ssdhot_ssd300_layout_host(perm.data_ptr(), 8732) == 0


def test_forward_heads_is_forward_without_the_tail():
    """ssdhot.forward_heads returns the twelve head outputs whose reference packing (SFS:249-269) is model(x), and
    tests/_util.unpack_heads inverts that packing -- all pure PyTorch, checked on the CPU."""
    import ssdhot
    torch.manual_seed(3)
    model = U.TinySSD().eval()
    x = torch.randn(2, 3, 300, 300)
    with torch.no_grad():
        loc_all, conf_all = model(x)
        loc_heads, conf_heads = ssdhot.forward_heads(model, x)
    assert [tuple(h.shape) for h in loc_heads] == [(2, a * 4, n, n) for n, a in U.LEVELS]
    assert [tuple(h.shape) for h in conf_heads] == [(2, a * 6, n, n) for n, a in U.LEVELS]
    back_loc, back_conf = O.pack_heads(loc_heads, conf_heads, 6)
    assert torch.equal(back_loc, loc_all) and torch.equal(back_conf, conf_all)
    for cl in (False, True):
        for h, u in zip(loc_heads + conf_heads, U.unpack_heads(loc_all, cl) + U.unpack_heads(conf_all, cl)):
            assert torch.equal(h, u)
            assert u.is_contiguous(memory_format=torch.channels_last if cl else torch.contiguous_format)
    with pytest.raises(ValueError):
        ssdhot.predict_heads(model, loc_heads[:5], conf_heads)           # validation precedes the CUDA requirement


@pytest.mark.skipif(not os.path.isfile("/root/reference/SSD_from_scratch.py"), reason="the reference tree is only mounted in the build container")
def test_forward_heads_against_the_real_reference_model():
    """With the unmodified reference class (random weights, CPU): mySSD.forward(x) == the reference packing of
    ssdhot.forward_heads(model, x), bit for bit -- forward_heads names the reference's own modules (SFS:236-262) correctly and
    leaves out nothing but the permute / cat tail (SFS:249-269)."""
    import importlib.util
    import ssdhot
    spec = importlib.util.spec_from_file_location("_ref_ssd_from_scratch", "/root/reference/SSD_from_scratch.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    torch.manual_seed(0)
    model = ref.mySSD({"a": 0, "b": 1, "c": 2, "d": 3, "e": 4}).eval()
    x = torch.randn(1, 3, 300, 300)
    with torch.no_grad():
        loc_all, conf_all = model(x)
        loc_heads, conf_heads = ssdhot.forward_heads(model, x)
    assert loc_all.shape == (1, 8732, 4) and conf_all.shape == (1, 8732, 6)
    back_loc, back_conf = O.pack_heads(loc_heads, conf_heads, 6)
    assert torch.equal(back_loc, loc_all) and torch.equal(back_conf, conf_all)
    assert torch.equal(model.priors, ssdhot.default_boxes())
