"""Shared-memory bank-conflict simulator for the NTT pass layouts (design aid, not product code).

Models 32 banks x 4 B; a warp access of W bytes/lane is split into phases of 128 B
(4B: 1 phase of 32 lanes, 8B: 2 phases of 16 lanes, 16B: 4 phases of 8 lanes); the cost of a
phase is the max number of distinct 128B-rows... i.e. distinct addresses mapping to one bank.
"""
import sys


def conflicts(addrs_bytes, width):
    lanes_per_phase = 128 // width
    worst = 0
    total = 0
    for p in range(0, 32, lanes_per_phase):
        banks = {}
        for a in addrs_bytes[p:p + lanes_per_phase]:
            for w in range(width // 4):
                b = ((a // 4) + w) % 32
                banks.setdefault(b, set()).add((a // 4) + w)
        deg = max(len(v) for v in banks.values())
        worst = max(worst, deg)
        total += deg
    return worst, total


def plan(logn, loge):
    npass = (logn + loge - 1) // loge
    first = logn - (npass - 1) * loge
    passes = []
    s = 0
    for p in range(npass):
        ns = first if p == 0 else loge
        fb = logn - loge if p == 0 else logn - (s + ns)
        passes.append((fb, s, ns))
        s += ns
    return passes


def swz(idx, sw_dst, sw_src):
    return idx ^ (((idx >> sw_src) & 7) << sw_dst)


def check(bits, logn, loge, sw_src=None, verbose=True):
    wbytes = bits // 8
    sw_dst = 1 if bits == 64 else 2
    R = max(loge - sw_dst, 0)
    if sw_src is None:
        sw_src = sw_dst + max(3, R)
    n, e = 1 << logn, 1 << loge
    tpp = n // e
    res = []
    for (fb, s0, ns) in plan(logn, loge):
        worst_all = 0
        cyc = 0
        for warp in range(max(1, tpp // 32)):
            for j in range(e):
                addrs = []
                for lane in range(min(32, tpp)):
                    t = warp * 32 + lane
                    low = t & ((1 << fb) - 1)
                    high = t >> fb
                    idx = (high << (fb + loge)) | (j << fb) | low
                    addrs.append(swz(idx, sw_dst, sw_src) * wbytes)
                while len(addrs) < 32:
                    addrs.append(addrs[-1])
                w, tot = conflicts(addrs, wbytes)
                worst_all = max(worst_all, w)
                cyc += tot
        # vector (16B) access for the fb=0 pass
        vec = None
        if fb == 0:
            cw = 16 // wbytes
            worst_v = 0
            for warp in range(max(1, tpp // 32)):
                for c in range(e // cw):
                    addrs = []
                    for lane in range(min(32, tpp)):
                        t = warp * 32 + lane
                        idx = t * e + c * cw
                        addrs.append(swz(idx, sw_dst, sw_src) * wbytes)
                    while len(addrs) < 32:
                        addrs.append(addrs[-1])
                    w, _ = conflicts(addrs, 16)
                    worst_v = max(worst_v, w)
            vec = worst_v
        res.append((fb, s0, ns, worst_all, vec))
    # coalesced copy in/out with 16B vectors
    cw = 16 // wbytes
    worst_c = 0
    for v0 in range(0, n // cw, 32):
        addrs = [swz((v0 + l) * cw, sw_dst, sw_src) * wbytes for l in range(32)]
        w, _ = conflicts(addrs, 16)
        worst_c = max(worst_c, w)
    if verbose:
        print(f"u{bits} logn={logn} loge={loge} tpp={tpp} sw_src={sw_src}: passes(fb,s0,ns,scalar-worst,vec16-worst)={res} copy16-worst={worst_c}")
    return res, worst_c


if __name__ == "__main__":
    for bits, logn, loge in [(64, 12, 4), (64, 13, 4), (64, 13, 5), (64, 14, 5), (64, 14, 4), (64, 11, 4), (64, 11, 3), (64, 10, 3), (64, 10, 4), (64, 10, 5),
                             (32, 10, 3), (32, 10, 4), (32, 10, 5), (32, 11, 3), (32, 11, 4), (32, 12, 4), (32, 12, 3)]:
        check(bits, logn, loge)
