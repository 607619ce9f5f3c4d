"""clock64 timeline of CTA 0 of one fused DTP -> linear launch (eqf_fused_set_timeline): where each role waits.

usage: python tools/fused_timeline.py [config] [E] [group] [N] [dtp 1|2] > timeline.txt
Roles: 0 TMA producer (stamp = stage free, about to issue), 1 MMA issuer (B landed / A ready / committed), 2 transform
(raw tile ready / a_ready arrived), 3 epilogue (tile start / end), 4 DTP set 0 (row block: tables ready; own k-tile: start,
tile written), 5 table helper warps (row block: buffer free, tables built)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from equiformer_b200 import _lib, ops  # noqa: E402
from equiformer_b200.nets.graph_attention_transformer import DepthwiseTensorProduct  # noqa: E402

CONFIGS = {"qm9_l2": ("128x0e+64x1e+32x2e", "1x0e+1x1e+1x2e"), "md17_l3": ("128x0e+64x1e+64x2e+32x3e", "1x0e+1x1e+1x2e+1x3e"),
           "oc20_l1": ("256x0e+128x1e", "1x0e+1x1e")}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "qm9_l2"
    E = int(sys.argv[2]) if len(sys.argv) > 2 else 32560
    group = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    N = int(sys.argv[4]) if len(sys.argv) > 4 else 32
    which = int(sys.argv[5]) if len(sys.argv) > 5 else 1
    irreps, sh = CONFIGS[name]
    plan = DepthwiseTensorProduct(irreps, sh, irreps, internal_weights=False, bias=False).tp.plan
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    n_nodes = max(E // 14, 2)
    if which == 1:
        xs = [torch.randn(n_nodes, 2 * l + 1, m, device=dev, generator=g) for l, m in plan.in1_blocks]
        Bs = [torch.randn(n_nodes, 2 * l + 1, m, device=dev, generator=g) for l, m in plan.in1_blocks]
        dst = torch.sort(torch.randint(0, n_nodes, (E,), device=dev, generator=g)).values
        src = torch.randint(0, n_nodes, (E,), device=dev, generator=g)
        gather, w = (src, dst, Bs), torch.randn(E, plan.weight_numel, device=dev, generator=g)
        off = torch.randn(plan.weight_numel, device=dev, generator=g)
    else:
        xs = [torch.randn(E, 2 * l + 1, m, device=dev, generator=g) for l, m in plan.in1_blocks]
        gather, w, off = None, torch.randn(plan.weight_numel, device=dev, generator=g), None
    y = torch.randn(E, plan.d_y, device=dev, generator=g)
    l, _p, K = plan.out_groups[group]
    Wt = torch.randn(K, N, device=dev, generator=g) / K ** 0.5
    run = lambda: ops.dtp_linear_fwd_raw(plan, group, xs, y, w, Wt, gather=gather, w_offset=off)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    buf = torch.zeros(6 * 2048, dtype=torch.int64, device=dev)
    lib = _lib.load()
    lib.eqf_fused_set_timeline(buf.data_ptr())
    run()
    torch.cuda.synchronize()
    lib.eqf_fused_set_timeline(None)
    t = buf.cpu().view(6, 2048)
    t0 = int(t[t > 0].min())
    k_tiles = K // 32
    print(f"# {name} E={E} group={group} (l={l}, K={K}, {k_tiles} k-tiles) N={N} dtp{which}; cycles relative to the first stamp of CTA 0")
    names = ["tma", "mma", "transform", "epilogue", "dtp set0", "table helper"]
    for r in range(6):
        v = [int(x) - t0 for x in t[r].tolist() if x > 0]
        print(f"## role {r} {names[r]}: {len(v)} stamps, last {v[-1] if v else 0}")
        print("   first 60:", v[:60])
    # DTP set 0 (role 4): per row block one stamp (tables ready) then (start, end) per own k-tile; set 0 owns the even
    # k-tiles of the CTA's running count, so blocks alternate between ceil and floor of k_tiles / 2
    v = [int(x) - t0 for x in t[4].tolist() if x > 0]
    print("## dtp set0 per row block (first 6): t_tables_ready | own k-tiles (t_start, wait + math)")
    pos, it = 0, 0
    for b in range(6):
        own = sum(1 for kt in range(k_tiles) if (it + kt) % 2 == 0)
        it += k_tiles
        if pos + 1 + 2 * own > len(v):
            break
        seg = v[pos:pos + 1 + 2 * own]
        pos += 1 + 2 * own
        print(f"   block {b}: tables ready {seg[0]} | {[(seg[1 + 2 * i], seg[2 + 2 * i] - seg[1 + 2 * i]) for i in range(own)]}")
    h = [int(x) - t0 for x in t[5].tolist() if x > 0]
    print("## table helper per row block (first 8): (t_buffer_free, build)")
    print("   ", [(h[2 * i], h[2 * i + 1] - h[2 * i]) for i in range(min(8, len(h) // 2))])
    m = [int(x) - t0 for x in t[1].tolist() if x > 0]
    print("## mma per k-tile (first 24): (t_B_landed, wait_A, issue)")
    print("   ", [(m[3 * i], m[3 * i + 1] - m[3 * i], m[3 * i + 2] - m[3 * i + 1]) for i in range(min(24, len(m) // 3))])


if __name__ == "__main__":
    main()
