"""Plain-numpy reading of the wire form (test infrastructure)."""
import numpy as np


def unpack_wire(hb_or_arrays):
    """Plain-numpy reading of the wire form (include/slamdunk_b200.h, sd_wire): -> per slot (pos, ref, mapq, flags, xi, nm, nmp,
    lseq, cigar ops, 2-bit base codes, qualities) -- the format's second implementation, next to wire_expand_kernel."""
    w = hb_or_arrays if isinstance(hb_or_arrays, dict) else hb_or_arrays.wire_arrays()
    T = w["tiles"]
    nt = len(T) - 1
    qb = w["qual_bits"]
    out = []
    for t in range(nt):
        h = T[t]
        v = int(h["var_off"])
        fl = int(h["flags"])
        pos32 = refs = shape = None
        if fl & 1:
            pos32 = w["var"][v:v + 32].view(np.int32); v += 32
        if fl & 2:
            refs = w["var"][v:v + 16].view(np.uint16); v += 16
        if fl & 4:
            shape = w["var"][v:v + 32]; v += 32
        base = int(h["base_off"])
        cig_at = v
        for k in range(32):
            s = t * 32 + k
            m = int(w["mqf"][s])
            mapq, flags = m & 0xFF, m >> 8
            pad = bool(flags & 0x80)
            pos = int(pos32[k]) if pos32 is not None else int(h["base_pos"]) + int(w["pos16"][s])
            ref = int(refs[k]) if refs is not None else int(h["ref"])
            if shape is not None:
                lseq, ncig = int(shape[k]) & 0xFFFF, int(shape[k]) >> 16
                ops = [int(x) for x in w["var"][cig_at:cig_at + ncig]]
                cig_at += ncig
            else:
                lseq, ncig = (0, 0) if pad else (int(h["lseq"]), 1)
                ops = [] if pad else [lseq << 4]
            xi = float(w["xi_dict"][w["xi16"][s]]) if len(w["xi16"]) else float(w["xi32"][s])
            if pad:
                out.append(None)
                continue
            idx = base + np.arange(lseq)
            codes = (w["seq"][idx // 4] >> (2 * (idx % 4))) & 3
            if qb == 8:
                q = w["qual"][base:base + lseq]
            else:
                per = 8 // qb
                q = np.array(w["qual_dict"], dtype=np.uint8)[(w["qual"][idx // per] >> (qb * (idx % per))) & ((1 << qb) - 1)]
            base += lseq
            out.append(dict(pos=pos, ref=ref, mapq=mapq, flags=flags, xi=xi, nm=int(w["nm16"][s]), nmp=int(w["nmp16"][s]) if len(w["nmp16"]) else 0,
                            lseq=lseq, ops=ops, codes=codes.tolist(), qual=[int(x) for x in q]))
    return out, w
