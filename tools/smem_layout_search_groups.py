import itertools, sys
def halves(addrs, ws):
    return [addrs[:16], addrs[16:]] if ws == 8 else [addrs]
def excess(addrs, ws):
    nb = 16 if ws == 8 else 32
    tot = 0; ideal = 0
    for h in halves(addrs, ws):
        h = [x for x in h if x is not None]
        if not h: continue
        banks = {}
        for x in set(h): banks[x % nb] = banks.get(x % nb, 0) + 1
        tot += max(banks.values()); ideal += 1
    return tot - ideal, ideal
def evaluate(N, NC, CPG, PL, FQ, CCQ, CQ, ws, align):
    N2 = N*N; L = CPG*N2; NT = CPG*NC*N
    QFIELD = N*PL + FQ; QCOMP = 3*QFIELD + CCQ; QCELL = NC*QCOMP + CQ
    TPG = (L + 31)//32*32
    exc = tot = 0
    # column
    for w in range(TPG//32):
        lanes = range(w*32, w*32+32)
        a = [ (l//N2)*QCELL + l % N2 if l < L else None for l in lanes]
        e, i = excess(a, ws); exc += e*10; tot += i*10   # weight: many column accesses
    # plane tasks
    tpc = NC*N
    slot = 16 if (align and tpc <= 16 and ws == 8) else (32 if (align and tpc <= 32 and ws == 4) else None)
    def task_of_lane(l):
        if slot:
            cw, r = divmod(l, slot)
            if cw >= CPG or r >= tpc: return None
            return cw*tpc + r
        return l if l < NT else None
    nl = CPG*slot if slot else NT
    for w in range((nl + 31)//32):
        ts = [task_of_lane(l) for l in range(w*32, w*32+32)]
        a = [ (t//tpc)*QCELL + ((t//N) % NC)*QCOMP + (t % N)*PL if t is not None else None for t in ts]
        e, i = excess(a, ws); exc += e*8; tot += i*8
    return exc, tot
if __name__ == "__main__":
    for ws in (8, 4):
        nb = 16 if ws == 8 else 32
        for N, CPG in ((5,2),(6,3),(7,2),(4,2),(3,3)):
            N2 = N*N
            best = None
            for align in (0, 1):
                for PL in range(N2, N2+4):
                    for FQ in range(16):
                        for CCQ in range(16):
                            base = 3*(3*(N*PL+FQ)+CCQ)
                            CQ = (N2 - base) % nb   # column accesses bank-contiguous across cells
                            e, t = evaluate(N, 3, CPG, PL, FQ, CCQ, CQ, ws, align)
                            key = (e, PL, FQ+CCQ+CQ)
                            if best is None or key < best[0]: best = (key, dict(align=align, PL=PL, FQ=FQ, CCQ=CCQ, CQ=CQ), t)
            print(ws, N, CPG, best)

def evaluate2(N, NC, CPG, PL, FQ, CCQ, CQ, ws, order):
    """dense plane tasks with a permuted enumeration order: order = tuple of 'w','c','k' fastest-last"""
    N2 = N*N; L = CPG*N2; NT = CPG*NC*N
    QFIELD = N*PL + FQ; QCOMP = 3*QFIELD + CCQ; QCELL = NC*QCOMP + CQ
    dims = {'w': CPG, 'c': NC, 'k': N}
    exc = tot = 0
    def decode(t):
        idx = {}
        for name in reversed(order):  # last = fastest
            idx[name] = t % dims[name]; t //= dims[name]
        return idx
    for w in range((NT + 31)//32):
        a = []
        for l in range(w*32, w*32+32):
            if l < NT:
                i = decode(l); a.append(i['w']*QCELL + i['c']*QCOMP + i['k']*PL)
            else: a.append(None)
        e, i_ = excess(a, ws); exc += e; tot += i_
    return exc, tot
if __name__ == "__main__":
    print("--- permuted task orders, FP64")
    for ws in (8,4):
      nb = 16 if ws == 8 else 32
      for N, CPG in ((6,3),(7,2),(6,2),(7,3)):
        N2 = N*N; best = None
        for order in itertools.permutations('wck'):
            for PL in range(N2, N2+4):
                for FQ in range(16):
                    for CCQ in range(16):
                        base = 3*(3*(N*PL+FQ)+CCQ)
                        CQ = (N2 - base) % nb
                        e, t = evaluate2(N, 3, CPG, PL, FQ, CCQ, CQ, ws, order)
                        key = (e, PL, FQ+CCQ+CQ)
                        if best is None or key < best[0]: best = (key, ''.join(order), dict(PL=PL, FQ=FQ, CCQ=CCQ, CQ=CQ), t)
        print(ws, N, CPG, best)
    print("--- FP32 N=5 dense orders")
    for ws, N, CPG in ((4,5,2),):
        nb = 32; N2=N*N; best=None
        for order in itertools.permutations('wck'):
            for PL in range(N2, N2+4):
                for FQ in range(16):
                    for CCQ in range(16):
                        base = 3*(3*(N*PL+FQ)+CCQ)
                        CQ = (N2 - base) % nb
                        e, t = evaluate2(N, 3, CPG, PL, FQ, CCQ, CQ, ws, order)
                        key = (e, PL, FQ+CCQ+CQ)
                        if best is None or key < best[0]: best = (key, ''.join(order), dict(PL=PL, FQ=FQ, CCQ=CCQ, CQ=CQ), t)
        print(ws, N, CPG, best)
