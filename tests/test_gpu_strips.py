"""Strip-sharded paths exercised on ONE GPU against the ORACLE (not against the unsharded CUDA result):
 * N logical strips writing into a device-resident panorama take the same code the N-GPU run takes - with
   gather_copy = False the level-0 blend kernel's staged 16-byte store variant (`blend_cell_kernel<2, 0>`, the one that
   stores over NVLink into rank 0's panorama), with gather_copy = True the local double-buffered strip + copy-engine push;
 * host-resident sources are uploaded band-wise (src_band_kernel): only the rows a strip can read;
 * the asymmetric halo (4 cells above / 3 below a strip) and the trimmed coarse blend rows must leave every owned row exact.
tests/test_gpu_multi.py runs the same thing across two processes / two GPUs when the box has them."""
import numpy as np
import pytest

import image_stitching_b200 as isb
from conftest import make_case, seam_masks_oracle
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


_ORACLE_CACHE = {}


def _oracle(case, key=None):
    if key is not None and key in _ORACLE_CACHE:
        return _ORACLE_CACHE[key]
    rig, imgs, gains, nb = case
    ref = orc.compose(imgs, rig.Ks, rig.Rs, rig.scale, rig.warp, nb, gains, seam_masks_oracle(rig))
    if key is not None:
        _ORACLE_CACHE[key] = ref
    return ref


CASES = {
    "cfg2_nb5": lambda: make_case("cfg2", 8, 5),      # 32-px cells: the cell kernel + staged stores at level 0
    "cfg3_nb3": lambda: make_case("cfg3", 16, 3),     # quad kernel at every level
    "cfg4_nb5": lambda: make_case("cfg4", 4, 5),      # cylindrical
    # small frames under a deep pyramid: most of every tile is REFLECT padding (mirror_pad_kernel), and the strip cuts clip tiles
    # whose padding reflects onto rows the strip does not hold (those tiles must compute their padding themselves)
    "cfg5_nb6": lambda: make_case("cfg5", 8, 6, max_images=40),
}


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("strips", [2, 3, 8])
@pytest.mark.parametrize("gather_copy", [False, True])
def test_logical_strips_device_panorama_vs_oracle(name, strips, gather_copy):
    case = CASES[name]()
    rig, imgs, gains, nb = case
    ref = _oracle(case, name)
    seams = seam_masks_oracle(rig)
    h, w = ref["mask"].shape
    dev = torch.device("cuda", 0)
    # 128-byte padded rows, as the N-GPU bench allocates them, and a pitch that is NOT a multiple of 16
    for p8, pm in (((w * 3 + 127) // 128 * 128, (w + 127) // 128 * 128), (w * 3 + 2, w + 6)):
        pano = torch.full((h, p8), 7, dtype=torch.uint8, device=dev)
        mask = torch.full((h, pm), 7, dtype=torch.uint8, device=dev)
        dimgs = [torch.from_numpy(im).to(dev) for im in imgs]
        rows = []
        for i in range(strips):
            c = isb.Composer(rig.warp, rig.scale, nb, strip_index=i, strip_count=strips, gather_copy=gather_copy)
            c.plan(isb.cameras_from_KR(rig.Ks, rig.Rs), [(rig.W, rig.H)] * rig.n)
            # device sources for even strips, host sources (band upload) for odd ones
            src = dimgs if i % 2 == 0 else imgs
            for _ in range(3):  # three runs: both staging slots of the copy path get reused
                r = c.run(src, gains, seams, out=pano, out_mask=mask, out_pitch=p8, mask_pitch=pm)
            c.sync()
            torch.cuda.synchronize()
            rows.append(r["strip_rows"])
            if i % 2 == 1 and strips >= 3:
                lo_hi = [c.source_band(k) for k in range(rig.n)]
                assert all(0 <= lo <= hi < rig.H for lo, hi in lo_hi)
                assert c.last_h2d_bytes() <= sum(im.nbytes for im in imgs)
        assert rows[0][0] == 0 and rows[-1][1] == h and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
        o8 = pano.cpu().numpy()[:, : w * 3].reshape(h, w, 3)
        om = mask.cpu().numpy()[:, :w]
        assert np.array_equal(om, ref["mask"])
        assert np.array_equal(o8, ref["result8"])
        # nothing outside the panorama's own bytes was touched (row padding)
        assert (pano.cpu().numpy()[:, w * 3:] == 7).all() and (mask.cpu().numpy()[:, w:] == 7).all()


def test_band_upload_matches_full_upload_and_oracle():
    """Host sources, 4 strips, host output: every strip uploads its row band only and the assembled panorama is the oracle's."""
    case = make_case("cfg2", 8, 4)
    rig, imgs, gains, nb = case
    ref = _oracle(case)
    seams = seam_masks_oracle(rig)
    h, w = ref["mask"].shape
    out8, outm, out16 = np.zeros((h, w, 3), np.uint8), np.zeros((h, w), np.uint8), np.zeros((h, w, 3), np.int16)
    total = 0
    for i in range(4):
        c = isb.Composer(rig.warp, rig.scale, nb, strip_index=i, strip_count=4)
        c.plan(isb.cameras_from_KR(rig.Ks, rig.Rs), [(rig.W, rig.H)] * rig.n)
        c.run(imgs, gains, seams, out=out8, out_mask=outm, out16=out16)
        total += c.last_h2d_bytes()
    assert np.array_equal(outm, ref["mask"]) and np.array_equal(out16, ref["result16"]) and np.array_equal(out8, ref["result8"])
    assert total < 4 * sum(im.nbytes for im in imgs)  # less than four full uploads


def test_blender_reuse_does_not_leak_device_memory():
    """prepare / feed x n / blend cycles on one MultiBandBlender: the per-feed tile allocations are released by the next
    prepare() (PyramidEngine::reset kept them)."""
    rng = np.random.default_rng(5)
    b = isb.MultiBandBlender(0, 3)
    corners, sizes = [(0, 0), (300, 10), (620, -5)], [(400, 300)] * 3
    imgs = [rng.integers(0, 256, (300, 400, 3)).astype(np.int16) for _ in range(3)]
    masks = [np.full((300, 400), 255, np.uint8) for _ in range(3)]

    def cycle():
        b.prepare(corners, sizes)
        for im, m, c in zip(imgs, masks, corners):
            b.feed(im, m, c)
        return b.blend()

    first = cycle()
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    for _ in range(6):
        out = cycle()
    torch.cuda.synchronize()
    free1 = torch.cuda.mem_get_info()[0]
    assert np.array_equal(out[0], first[0]) and np.array_equal(out[1], first[1])
    assert free0 - free1 < 8 << 20, f"device memory shrank by {(free0 - free1) >> 20} MiB over 6 cycles"


@pytest.mark.parametrize("depth", [2, 3])
def test_pipeline_depth_runs_in_flight_are_exact(depth):
    """isb_config.pipeline_depth: consecutive runs are served by independent pyramid sets on internal streams; every panorama
    is the oracle's, outputs are valid behind join() on the caller's stream, and new pixels per run are honoured."""
    case = make_case("cfg2", 8, 5)
    rig, imgs, gains, nb = case
    seams = seam_masks_oracle(rig)
    ref = _oracle(case)
    imgs2 = [np.ascontiguousarray(im[::-1]) for im in imgs]  # a second frame set
    ref2 = orc.compose(imgs2, rig.Ks, rig.Rs, rig.scale, rig.warp, nb, gains, seams)
    h, w = ref["mask"].shape
    dev = torch.device("cuda", 0)
    st = torch.cuda.Stream(device=dev)
    isb.set_stream(st.cuda_stream)
    try:
        with torch.cuda.stream(st):
            d1 = [torch.from_numpy(im).to(dev) for im in imgs]
            d2 = [torch.from_numpy(im).to(dev) for im in imgs2]
            c = isb.Composer(rig.warp, rig.scale, nb, pipeline_depth=depth)
            c.plan(isb.cameras_from_KR(rig.Ks, rig.Rs), [(rig.W, rig.H)] * rig.n)
            outs = [(torch.zeros((h, w, 3), dtype=torch.uint8, device=dev), torch.zeros((h, w), dtype=torch.uint8, device=dev))
                    for _ in range(2 * depth)]
            for k, (o, m) in enumerate(outs):
                c.run(d1 if k % 2 == 0 else d2, gains, seams, out=o, out_mask=m)
            c.join()
            got = [(o.clone(), m.clone()) for o, m in outs]  # enqueued on the caller's stream, behind the join
        st.synchronize()
        for k, (o, m) in enumerate(got):
            r = ref if k % 2 == 0 else ref2
            assert np.array_equal(m.cpu().numpy(), r["mask"]) and np.array_equal(o.cpu().numpy(), r["result8"]), k
        # host outputs: valid after sync()
        ho, hm = np.zeros((h, w, 3), np.uint8), np.zeros((h, w), np.uint8)
        c.run(imgs2, gains, seams, out=ho, out_mask=hm)
        c.sync()
        assert np.array_equal(hm, ref2["mask"]) and np.array_equal(ho, ref2["result8"])
    finally:
        isb.set_stream(None)
