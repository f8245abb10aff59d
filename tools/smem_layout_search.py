"""Offline search for the shared-memory strides of the cell kernel
(csrc/cell_kernel.cuh): counts bank-conflict passes of every warp-level access
pattern of the 7 phases for a layout  idx = cell*CELL + k*KS + j*JS + i.
FP64: a warp request is served per half-warp over 16 8-byte banks;
FP32: full warp over 32 4-byte banks.  Same address = broadcast."""
import itertools
import sys


def passes(addrs, wordsize):
    if wordsize == 8:
        halves = [addrs[:16], addrs[16:]]
        nb = 16
    else:
        halves = [addrs]
        nb = 32
    tot = 0
    for h in halves:
        h = [x for x in h if x is not None]
        if not h:
            continue
        banks = {}
        for x in set(h):
            banks.setdefault(x % nb, 0)
            banks[x % nb] += 1
        tot += max(banks.values())
    return tot


def ideal(lanes, wordsize):
    if wordsize == 8:
        return (1 if lanes > 0 else 0) + (1 if lanes > 16 else 0)
    return 1


def evaluate(N, JS, KS, CELL, wordsize, lanes_total=32):
    N2 = N * N
    cpw = max(1, lanes_total // N2)
    nl = min(lanes_total, cpw * N2) if N2 <= 32 else 32
    cost = 0
    base = 0
    pats = {
        "ab_JS": lambda a, b, x: x * KS + b * JS + a,
        "ab_KS": lambda a, b, x: b * KS + x * JS + a,
        "aJS_bKS": lambda a, b, x: b * KS + a * JS + x,
        "bc_y": lambda a, b, x: b * KS + x * JS + 0,
        "bc_z": lambda a, b, x: x * KS + a * JS + 0,
    }
    weights = {"ab_JS": 2, "ab_KS": 4, "aJS_bKS": 6, "bc_y": 2 * N, "bc_z": 2 * N}
    detail = {}
    for name, f in pats.items():
        worst = 0
        for x in range(N):
            # for N2 > 32 look at each warp of the cell
            if N2 <= 32:
                warps = [list(range(32))]
            else:
                nw = (N2 + 31) // 32
                warps = [list(range(w * 32, w * 32 + 32)) for w in range(nw)]
            for wl in warps:
                addrs = []
                for lane in wl:
                    if N2 <= 32:
                        cw, t = divmod(lane, N2)
                        if cw >= cpw:
                            addrs.append(None)
                            continue
                    else:
                        cw, t = 0, lane
                        if t >= N2:
                            addrs.append(None)
                            continue
                    a, b = t % N, t // N
                    addrs.append(cw * CELL + f(a, b, x))
                nact = sum(1 for q in addrs if q is not None)
                worst += passes(addrs, wordsize) - ideal(nact, wordsize)
        detail[name] = worst
        cost += weights[name] * worst
    return cost, detail


def search(N, wordsize, ncomp=3):
    best = None
    for JS in (N, N + 1, N + 2):
        for kpad in range(0, 6):
            KS = N * JS + kpad
            cstr = N * KS  # one comp buffer
            for cpad in range(0, 17):
                CELL = 3 * ncomp * cstr + cpad  # 3 buffers x ncomp comps
                c, d = evaluate(N, JS, KS, CELL, wordsize)
                key = (c, CELL)
                if best is None or key < best[0]:
                    best = (key, JS, KS, CELL, d)
    return best


if __name__ == "__main__":
    for ws in (8, 4):
        for ncomp in (3, 1):
            for N in range(2, 8):
                (c, _), JS, KS, CELL, d = search(N, ws, ncomp)
                print(f"word={ws} ncomp={ncomp} N={N}: JS={JS} KS={KS} CELL={CELL} extra_passes={c} {d}")
