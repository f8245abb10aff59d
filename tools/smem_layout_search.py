"""Offline check of the shared-memory strides of the plane/column cell kernel (csrc/cell_kernel.cuh, v3):
nested strides: plane PL; bufZ component N*PL + CCZ, cell NC*comp + CZ; bufQ field N*PL + FQ, component 3*field + CCQ,
cell NC*comp + CQ; element (j*N + i).  FP64: a warp request is served per half-warp over 16 8-byte banks; FP32: 32 4-byte banks.
Prints the wavefront excess of every access pattern for the chosen pads."""
import sys


def passes(addrs, wordsize):
    halves = [addrs[:16], addrs[16:]] if wordsize == 8 else [addrs]
    nb = 16 if wordsize == 8 else 32
    tot = 0
    for h in halves:
        h = [x for x in h if x is not None]
        if not h:
            continue
        banks = {}
        for x in set(h):
            banks[x % nb] = banks.get(x % nb, 0) + 1
        tot += max(banks.values())
    return tot


def ideal(addrs, wordsize):
    halves = [addrs[:16], addrs[16:]] if wordsize == 8 else [addrs]
    return sum(1 for h in halves if any(x is not None for x in h))


def group_cells(N, NC):
    """mirror of group_cells() in csrc/cell_kernel.cuh"""
    if NC == 1:
        return {2: 8, 3: 7, 4: 8, 5: 6, 6: 4}.get(N, 2)
    return 32 // (N * N) if N * N <= 16 else {5: 2, 6: 3, 7: 2}[N]


def evaluate(N, NC, PL, CZ, CQ, ws, CCZ=0, CCQ=0, FQ=0, which="zq"):
    N2 = N * N
    CPG = group_cells(N, NC)
    L = CPG * N2
    NT = CPG * NC * N
    exc = 0
    tot = 0
    nwarps = (L + 31) // 32

    def zaddr(cw, c, kq, e):
        return cw * (NC * (N * PL + CCZ) + CZ) + c * (N * PL + CCZ) + kq * PL + e

    def qaddr(cw, c, f, kq, e):
        return cw * (NC * (3 * (N * PL + FQ) + CCQ) + CQ) + c * (3 * (N * PL + FQ) + CCQ) + f * (N * PL + FQ) + kq * PL + e

    for w in range(nwarps):
        lanes = range(w * 32, w * 32 + 32)
        # column threads: fixed (c, kq[, f]) ; lane -> (cw, ji)
        for c in range(NC):
            for kq in range(N):
                a = [zaddr(l // N2, c, kq, l % N2) if l < L else None for l in lanes]
                if "z" in which:
                    exc += passes(a, ws) - ideal(a, ws); tot += ideal(a, ws)
                for f in range(3 if "q" in which else 0):
                    a = [qaddr(l // N2, c, f, kq, l % N2) if l < L else None for l in lanes]
                    exc += 2 * (passes(a, ws) - ideal(a, ws)); tot += 2 * ideal(a, ws)
        # plane tasks: fixed element e ; lane t -> (cw, c, kq)
        for t0 in range(0, NT, 32 * nwarps):
            ts = [t0 + l if t0 + l < NT else None for l in lanes]
            for e in range(N2):
                a = [zaddr(t // (NC * N), (t // N) % NC, t % N, e) if t is not None else None for t in ts]
                if "z" in which:
                    exc += 2 * (passes(a, ws) - ideal(a, ws)); tot += 2 * ideal(a, ws)
                for f in range(3 if "q" in which else 0):
                    a = [qaddr(t // (NC * N), (t // N) % NC, f, t % N, e) if t is not None else None for t in ts]
                    exc += 2 * (passes(a, ws) - ideal(a, ws)); tot += 2 * ideal(a, ws)
    return exc, tot


if __name__ == "__main__":
    # usage: smem_layout_search.py [max extra plane stride] [N ...]
    R = range(0, 16, 1)
    extra = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    only = [int(a) for a in sys.argv[2:]]
    for ws in (8, 4):
        for NC in (3, 1):
            for N in range(2, 8):
                if only and N not in only:
                    continue
                N2 = N * N
                best = None
                for PL in range(N2, N2 + extra):
                    bz = min(((evaluate(N, NC, PL, CZ, 0, ws, CCZ, 0, 0, "z")[0], CZ + CCZ, CZ, CCZ) for CZ in R for CCZ in (R if NC > 1 else [0])))
                    bq = min(((evaluate(N, NC, PL, 0, CQ, ws, 0, CCQ, FQ, "q")[0], CQ + CCQ + FQ, CQ, CCQ, FQ) for CQ in R for CCQ in (R if NC > 1 else [0]) for FQ in R))
                    key = (bz[0] + bq[0], PL)
                    if best is None or key < best[0]:
                        best = (key, PL, bz, bq)
                t = evaluate(N, NC, best[1], 0, 0, ws)[1]
                print(f"B2_PAD3({ws // 4}, {NC}, {N}, /*PL*/ {best[1]}, /*CZ*/ {best[2][2]}, /*CCZ*/ {best[2][3]}, /*CQ*/ {best[3][2]}, /*CCQ*/ {best[3][3]}, /*FQ*/ {best[3][4]})  // excess {best[0][0]} of {t}", flush=True)
