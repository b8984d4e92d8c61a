"""Per-tile statistics of one view of the headline scene from the CPU oracle (dev tool; test infrastructure only):

  python tools/tile_stats.py lists <view>   tile-list lengths, and for the longest tiles the contributors per pixel and per
                                           8x4 block -- what bounds a one-frame launch
  python tools/tile_stats.py merge <view>   how many (pixel, Gaussian) pairs of a warp hit the same Gaussian in the same
                                           trip of the backward walk -- the reductions a __match_any_sync merge could
                                           save
view: 0..7 = orbit view, -1 = the canonical camera."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import numpy as np, cases, oracle_cpu

mode = sys.argv[1] if len(sys.argv) > 1 else "lists"
view = int(sys.argv[2]) if len(sys.argv) > 2 else 1
c = cases.f3d_case(0, 256, 256, view if view >= 0 else None)
cn = oracle_cpu.case_to_numpy(c)
pre = oracle_cpu.preprocess(cn)
b = oracle_cpu.binning(256, 256, pre["means2D"], pre["depths"], pre["radii"], pre["tiles_touched"])
W = H = 256; fx = W / (2 * cn["tanfovx"]); fy = H / (2 * cn["tanfovy"])
v2g = pre["view2gaussian"].astype(np.float64); wop = pre["conic_opacity"][:, 3].astype(np.float64)
rng = b["ranges"].astype(np.int64); pl = b["point_list"].astype(np.int64)
lens = rng[:, 1] - rng[:, 0]


def blended(tile):
    """[256 pixels, n records] bool: the pairs the forward blends (contributing and before saturation)."""
    ty, tx = divmod(tile, 16)
    lo, hi = rng[tile]; n = hi - lo
    q = v2g[pl[lo:hi]]; w = wop[pl[lo:hi]]
    px = tx * 16 + np.arange(16); py = ty * 16 + np.arange(16)
    RX = ((px + 0.5 - W / 2.) / fx)[None, :].repeat(16, 0).reshape(-1)[:, None]
    RY = ((py + 0.5 - H / 2.) / fy)[:, None].repeat(16, 1).reshape(-1)[:, None]
    n0 = q[None, :, 0] * RX + q[None, :, 1] * RY + q[None, :, 2]; n1 = q[None, :, 1] * RX + q[None, :, 3] * RY + q[None, :, 4]
    n2 = q[None, :, 2] * RX + q[None, :, 4] * RY + q[None, :, 5]
    AA = n0 * RX + n1 * RY + n2; BB = 2 * (q[None, :, 6] * RX + q[None, :, 7] * RY + q[None, :, 8]); CC = q[None, :, 9]
    t = -BB / (2 * AA); mv = -(BB / AA) * (BB / 4) + CC
    alpha = np.minimum(0.99, w[None, :] * np.exp(np.minimum(0, -0.5 * mv)))
    contrib = (t > 0.2) & (alpha >= 1 / 255.)
    T = np.cumprod(1 - np.where(contrib, alpha, 0.0), axis=1)
    stop = T < 1e-4
    first_stop = np.where(stop.any(1), stop.argmax(1), n)
    return contrib & (np.arange(n)[None, :] < first_stop[:, None]), first_stop


if mode == "lists":
    print(f"view {view}: R {lens.sum()}, tile list mean {lens.mean():.0f} max {lens.max()}, percentiles 10/50/90/99 "
          f"{np.percentile(lens, [10, 50, 90, 99])}")
    order = np.argsort(-lens)
    for tile in list(order[:6]) + list(order[100:102]):
        work, first_stop = blended(tile)
        per_pix = work.sum(1).reshape(16, 16)
        wm = [int(per_pix[wy * 4:(wy + 1) * 4, wx * 8:(wx + 1) * 8].max()) for wy in range(4) for wx in range(2)]
        print(f"tile {tile}: {lens[tile]} records ({(lens[tile] + 127) // 128} chunks), saturated pixels {(first_stop < lens[tile]).sum()}, "
              f"contributors per pixel mean {per_pix.mean():.1f} max {per_pix.max()}, heaviest pixel per 8x4 block {wm}")
else:
    rs = np.random.RandomState(0)
    pairs = groups = trips = 0
    for tile in rs.choice(256, 24, replace=False):
        if lens[tile] <= 0:
            continue
        work = blended(tile)[0].reshape(16, 16, -1)
        for wy in range(4):
            for wx in range(2):
                blk = work[wy * 4:(wy + 1) * 4, wx * 8:(wx + 1) * 8].reshape(32, -1)
                lists = [np.nonzero(blk[l])[0][::-1] for l in range(32)]          # back to front
                for k in range(max(len(x) for x in lists)):
                    cur = [x[k] for x in lists if len(x) > k]
                    pairs += len(cur); groups += len(set(cur)); trips += 1
    print(f"view {view}: {pairs} pairs in {trips} warp trips, {groups} distinct (warp trip, Gaussian) groups -> same-trip merging "
          f"would save {1 - groups / pairs:.1%} of the reductions; {pairs / trips:.1f} active lanes and {groups / trips:.1f} Gaussians per trip")
