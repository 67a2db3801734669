"""TEST INFRASTRUCTURE (like the rest of oracle/): a plain-Python restatement of the
reference's LimEx-32 runtime in block mode, read straight from the engine's bytes
(struct NFA + struct LimExNFA32, src/nfa/limex_internal.h:102-203).  Only tests/ may
import it; the product never does.

Follows, for one block scanned the way Rose runs an outfix (queue {START@0, TOP@0,
END@len} through nfaExecLimEx32_Q, then nfaExecLimEx32_testEOD):
  * moNfaTop32: the top at offset 0 ORs `init` into the state (limex_common_impl.h:225-232)
  * LOOP_NOACCEL_FN / STREAM_FN (limex_runtime_impl.h:209-243, 246-366): per byte, limited
    shifts, then the exceptions of the states that are on, then succ & reach[reachMap[c]];
    after the last byte the accepts of the final state fire at offset len
  * processExceptional32 / RUN_EXCEPTION_FN (limex_exceptional.h:92-190, 190-330): in
    ascending state order, reports at the current offset unless this is the first byte
    of the scan (NO_OUTPUT | FIRST_BYTE), successors collected aside, squash applied to
    the shift successors for LIMEX_SQUASH_CYCLIC / _REPORT
  * moProcessAccepts32 (limex_common_impl.h:116-176) and moNfaTestEod32 (:192-218)
Bounded repeats and acceleration are not modelled (the emitters do not produce them)."""
import struct

INVALID = 0xffffffff
NFA_HDR = 64
O = {"reachMap": 0, "reachSize": 256, "acceptCount": 276, "acceptOffset": 280, "acceptEodCount": 284,
     "acceptEodOffset": 288, "exceptionCount": 292, "exceptionOffset": 296, "repeatCount": 300, "flags": 328,
     "init": 332, "initDS": 336, "accept": 340, "acceptAtEOD": 344, "exceptionMask": 368, "shift": 380,
     "shiftCount": 412, "shiftAmount": 416, "sizeof": 640}


def _u32(b, off):
    return struct.unpack_from("<I", b, off)[0]


def _reports(lx, off):
    out = []
    while True:
        r = _u32(lx, off)
        if r == INVALID:
            return out
        out.append(r)
        off += 4


def _accepts(lx, found, mask, table, to, out, block):
    for bit in range(32):
        if not (found >> bit) & 1:
            continue
        idx = bin(mask & ((1 << bit) - 1)).count("1")
        single, reports, _sq = struct.unpack_from("<B3xII", lx, table + 12 * idx)
        for r in ([reports] if single else _reports(lx, reports)):
            out.append((r, block, to))


def walk_blocks(engine, data, offsets, lengths):
    """[(report, block, to)] in callback order"""
    assert engine[8] == 0, "not LIMEX_NFA_32"
    lx = bytes(engine[NFA_HDR:])
    assert _u32(lx, O["repeatCount"]) == 0
    reach_map = lx[0:256]
    reach = [_u32(lx, O["sizeof"] + 4 * i) for i in range(_u32(lx, O["reachSize"]))]
    nshift = _u32(lx, O["shiftCount"])
    shifts = [(_u32(lx, O["shift"] + 4 * k), lx[O["shiftAmount"] + k]) for k in range(nshift)]
    emask = _u32(lx, O["exceptionMask"])
    eoff = _u32(lx, O["exceptionOffset"])
    exc = [struct.unpack_from("<IIIIBB", lx, eoff + 20 * i) for i in range(_u32(lx, O["exceptionCount"]))]
    accept, accept_eod = _u32(lx, O["accept"]), _u32(lx, O["acceptAtEOD"])
    init = _u32(lx, O["init"])
    out = []
    for b, (o, n) in enumerate(zip(offsets, lengths)):
        o, n = int(o), int(n)
        s = init
        for i in range(n):
            succ = 0
            for m, a in shifts:
                succ |= ((s & m) << a) & 0xffffffff
            est = s & emask
            if est:
                local = 0
                for bit in range(32):
                    if not (est >> bit) & 1:
                        continue
                    squash, successors, reports, _rep, has_squash, _trig = exc[bin(emask & ((1 << bit) - 1)).count("1")]
                    if reports != INVALID and i != 0:
                        out.extend((r, b, i) for r in _reports(lx, reports))
                    local |= successors
                    if has_squash in (1, 3):
                        succ &= squash
                succ |= local
            s = succ & reach[reach_map[int(data[o + i])]]
        if n and (s & accept):
            _accepts(lx, s & accept, accept, _u32(lx, O["acceptOffset"]), n, out, b)
        if _u32(lx, O["acceptEodCount"]) and (s & accept_eod):
            _accepts(lx, s & accept_eod, accept_eod, _u32(lx, O["acceptEodOffset"]), n, out, b)
    return out
