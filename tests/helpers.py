"""Test helpers: oracle outputs -> the flat result arrays of the C ABI (frs_result)."""
import re

import numpy as np

from oracle import segment_oracle as orc

_GAP = re.compile(r"^([0-9]+)-([0-9]+):([0-9]+)$")
_POLY = re.compile(r"^([SE])([AT])_([0-9]+):([0-9]+)$")
_CLIP = re.compile(r"^([SE])SC:([0-9]+)$")


def oracle_result_arrays(tints, prm):
    """Runs the oracle on (deep copies of) ``tints`` and returns (otints, arrays) where arrays has the
    layout of ``frs_result`` for a batch packed by ``freddie_b200.pack.pack_tints``."""
    import copy
    otints = copy.deepcopy(tints)
    tfo, fpos, tdo, digits, heads, goff, grec = [0], [], [0], [], [], [0], []
    for t in otints:
        it = orc.segment_tint(t, prm, keep=True)
        fpos.extend(t["final_positions"])
        tfo.append(len(fpos))
        for row in it["rows"]:
            digits.extend(48 + d for d in row)
        tdo.append(len(digits))
        for r in t["reads"]:
            h = [0] * 8
            recs = []
            if r["gaps"]:
                h[0] = 1
            for g in r["gaps"]:
                m = _GAP.match(g)
                if m:
                    recs.append([int(x) for x in m.groups()])
                    continue
                m = _POLY.match(g)
                if m:
                    kind = 1 if m.group(2) == "A" else 2
                    if m.group(1) == "S":
                        h[0] |= kind << 8
                        h[1], h[2] = int(m.group(3)), int(m.group(4))
                    else:
                        h[0] |= kind << 16
                        h[4], h[5] = int(m.group(3)), int(m.group(4))
                    continue
                m = _CLIP.match(g)
                assert m, g
                if m.group(1) == "S":
                    h[3] = int(m.group(2))
                else:
                    h[6] = int(m.group(2))
            recs.sort()
            heads.extend(h)
            for rc in recs:
                grec.extend(rc)
            goff.append(goff[-1] + len(recs))
    arrays = dict(
        tint_final_off=np.array(tfo, dtype=np.int32), final_pos=np.array(fpos, dtype=np.int32),
        tint_digit_off=np.array(tdo, dtype=np.int64), digits=np.array(digits, dtype=np.uint8),
        read_head=np.array(heads, dtype=np.int32), read_gap_off=np.array(goff, dtype=np.int32),
        gap_rec=np.array(grec, dtype=np.int32),
    )
    return otints, arrays
