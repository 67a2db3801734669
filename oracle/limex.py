"""TEST INFRASTRUCTURE (like the rest of oracle/): a plain-Python restatement of the
reference's LimEx runtime (32- to 512-state models) in block mode, read straight from the
engine's bytes (struct NFA + struct LimExNFA32 ... 512, src/nfa/limex_internal.h:102-203).  Only tests/ may
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
The models above 64 states shift each 64-bit lane of the state on its own (lshift_m128 = lshift64_m128 ...,
src/util/uniform_ops.h:142-145).
Bounded repeats and acceleration are not modelled (the emitters do not produce them)."""
import json
import os
import struct

INVALID = 0xffffffff
NFA_HDR = 64
# field offsets of struct LimExNFA32 / LimExNFA64 and their exception records: the reference's own
# sizeof / offsetof, recorded in tests/golden/ref_layout.json (tests/golden/gen_ref_layout.py)
with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                       "ref_layout.json")) as _f:
    _LAYOUT = json.load(_f)


def _offsets(bits):
    name = "LimExNFA%d" % bits
    o = {k.split(".", 1)[1]: v for k, v in _LAYOUT.items() if k.startswith(name + ".")}
    o["sizeof"] = _LAYOUT["sizeof(%s)" % name]
    o["exc_size"] = _LAYOUT["sizeof(NFAException%d)" % bits]
    return o


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
    for bit in range(found.bit_length()):
        if not (found >> bit) & 1:
            continue
        idx = bin(mask & ((1 << bit) - 1)).count("1")
        single, reports, _sq = struct.unpack_from("<B3xII", lx, table + 12 * idx)
        for r in ([reports] if single else _reports(lx, reports)):
            out.append((r, block, to))


def _state(b, off, bits):
    return int.from_bytes(b[off:off + bits // 8], "little")


def _shift_lanes(v, a, bits):
    """LSHIFT_STATE: a plain shift up to 64 bits, lane by lane (64-bit lanes) above"""
    if bits <= 64:
        return (v << a) & ((1 << bits) - 1)
    m64 = (1 << 64) - 1
    return sum(((((v >> (64 * j)) & m64) << a) & m64) << (64 * j) for j in range(bits // 64))


def walk_blocks(engine, data, offsets, lengths):
    """[(report, block, to)] in callback order"""
    assert engine[8] in (0, 1, 2, 3, 5), "not one of the LimEx models restated here"
    bits = {0: 32, 1: 64, 2: 128, 3: 256, 5: 512}[engine[8]]
    O = _offsets(bits)
    lx = bytes(engine[NFA_HDR:])
    assert _u32(lx, O["repeatCount"]) == 0
    reach_map = lx[0:256]
    reach = [_state(lx, O["sizeof"] + (bits // 8) * i, bits) for i in range(_u32(lx, O["reachSize"]))]
    nshift = _u32(lx, O["shiftCount"])
    shifts = [(_state(lx, O["shift"] + (bits // 8) * k, bits), lx[O["shiftAmount"] + k]) for k in range(nshift)]
    emask = _state(lx, O["exceptionMask"], bits)
    eoff = _u32(lx, O["exceptionOffset"])
    sb = bits // 8

    def exception(off):
        return (_state(lx, off, bits), _state(lx, off + sb, bits)) + struct.unpack_from("<IIBB", lx, off + 2 * sb)
    exc = [exception(eoff + O["exc_size"] * i) for i in range(_u32(lx, O["exceptionCount"]))]
    accept, accept_eod = _state(lx, O["accept"], bits), _state(lx, O["acceptAtEOD"], bits)
    init = _state(lx, O["init"], bits)
    out = []
    for b, (o, n) in enumerate(zip(offsets, lengths)):
        o, n = int(o), int(n)
        s = init
        for i in range(n):
            succ = 0
            for m, a in shifts:
                succ |= _shift_lanes(s & m, a, bits)
            est = s & emask
            if est:
                local = 0
                for bit in range(est.bit_length()):
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
