"""LZO1X block compressor for the PV writer (host-side plumbing, no compute on the hot path).

The reference compresses every frame payload with minilzo's lzo1x_1_compress and keeps the compressed form when it is smaller
(Frame::serialize, Application/src/ProcessedVideo/pv.cpp:726-763); readers decode with lzo1x_decompress (pv.cpp:315-336).
minilzo is a third-party library vendored by the reference (Application/src/ProcessedVideo/lzo/, LZO 2.10); this module is a
clean-room encoder for the published LZO1X stream format -- a greedy hash matcher, not minilzo's match finder, so the bytes
differ from what TRex would write while every LZO1X decoder reads them.  tests/test_pv_writer.py round-trips its output
through the reference's own decompressor.

Stream format (as the decoder consumes it):
  first byte 18..255          initial literal run of (byte - 17) bytes
  t < 16 (after a match with no trailing literals, or at the start)   literal run of t + 3 bytes; t == 0: zero bytes add 255
                              each, the first non-zero byte b ends the count: length = 18 + 255 * zeros + b
  t >= 64   M2  length ((t >> 5) - 1) + 2 = 3..8, distance 1 + ((t >> 2) & 7) + (next << 3) <= 2048
  32..63    M3  length (t & 31) + 2 (0: extended, base 33), then u16 LE: distance 1 + (v >> 2) <= 16384
  16..31    M4  length (t & 7) + 2 (0: extended, base 9), distance 16384 + ((t & 8) << 11) + (v >> 2); distance 16384 with
                t == 17 and v == 0 is the end-of-stream marker
  the two low bits of the last distance byte (M2: of t) give 0..3 literals that follow the match directly; the next
  instruction is then a match again.
"""
from __future__ import annotations

M2_MAX_LEN, M2_MAX_OFF = 8, 0x800
M3_MAX_OFF, M4_MAX_OFF = 0x4000, 0xBFFF


def _ext(out: bytearray, n: int):
    """Extended length: n > 0 as zero bytes worth 255 each and a final byte 1..255."""
    while n > 255:
        out.append(0); n -= 255
    out.append(n)


def compress(data: bytes) -> bytes:
    src = bytes(data)
    n = len(src)
    out = bytearray()
    table: dict[bytes, int] = {}
    lit_start = 0                 # pending literals: src[lit_start:ip]
    patch = -1                    # index in `out` of the byte whose two low bits may carry 1..3 trailing literals
    first = True
    ip = 0

    def flush_literals(end: int):
        nonlocal lit_start, patch, first
        k = end - lit_start
        if k == 0:
            return
        if first and k <= 238:
            out.append(17 + k)
        elif patch >= 0 and k <= 3:
            out[patch] |= k
        elif k <= 3:
            # 1..3 literals cannot follow a literal-less position in the stream: only reachable at the very start with k <= 3,
            # which the first branch takes
            raise AssertionError("short literal run without a preceding match")
        elif k <= 18:
            out.append(k - 3)
        else:
            out.append(0); _ext(out, k - 18)
        out.extend(src[lit_start:end])
        lit_start = end
        first = False
        patch = -1

    while ip + 3 <= n:
        key = src[ip:ip + 3]
        cand = table.get(key, -1)
        table[key] = ip
        if cand >= 0 and ip - cand <= M4_MAX_OFF:
            # extend the match
            m = 3
            limit = n - ip
            while m < limit and src[cand + m] == src[ip + m]:
                m += 1
            off = ip - cand
            pending = ip - lit_start
            if 0 < pending <= 3 and patch < 0 and not first:
                # 1..3 literals after an explicit literal run would need a match in between: cannot happen (runs are merged)
                raise AssertionError
            if first and pending == 0:
                # a stream cannot start with a match; not reachable (no history at ip = 0)
                raise AssertionError
            flush_literals(ip)
            if m <= M2_MAX_LEN and off <= M2_MAX_OFF:
                o = off - 1
                out.append(((m - 1) << 5) | ((o & 7) << 2)); patch = len(out) - 1
                out.append(o >> 3)
            elif off <= M3_MAX_OFF:
                o = off - 1
                if m <= 33:
                    out.append(32 | (m - 2))
                else:
                    out.append(32); _ext(out, m - 33)
                out.append((o << 2) & 0xFF); patch = len(out) - 1
                out.append((o >> 6) & 0xFF)
            else:
                o = off - 0x4000
                if m <= 9:
                    out.append(16 | ((o >> 11) & 8) | (m - 2))
                else:
                    out.append(16 | ((o >> 11) & 8)); _ext(out, m - 9)
                out.append((o << 2) & 0xFF); patch = len(out) - 1
                out.append((o >> 6) & 0xFF)
            # index a few positions inside the match so later data can refer to it
            for j in range(ip + 1, min(ip + m, n - 2), 1 if m < 64 else 7):
                table[src[j:j + 3]] = j
            ip += m
            lit_start = ip
        else:
            ip += 1
            # an explicit literal run directly after a match that already carries 1..3 trailing literals is not expressible:
            # trailing literals are only committed in flush_literals, so nothing to do here
    # tail
    if n - lit_start > 0:
        k = n - lit_start
        if not first and patch < 0 and k <= 3:
            # the previous instruction was an explicit literal run; merge is impossible now, so re-emit is needed.  Cannot
            # happen: literals are only flushed right before a match or here.
            raise AssertionError
        flush_literals(n)
    out += b"\x11\x00\x00"
    return bytes(out)


def decompress(block: bytes, max_out: int | None = None) -> bytes:
    """LZO1X block decoder (what lzo1x_decompress does for pv::Frame::read_from, pv.cpp:315-336).  Raises ValueError on a
    malformed stream or when the output would exceed max_out."""
    src = bytes(block)
    n = len(src)
    out = bytearray()
    ip = 0

    def copy_match(dist: int, length: int):
        if dist <= 0 or dist > len(out):
            raise ValueError("LZO1X: match distance outside the window")
        start = len(out) - dist
        if dist >= length:
            out.extend(out[start:start + length])
        else:
            for k in range(length):
                out.append(out[start + k])
        if max_out is not None and len(out) > max_out:
            raise ValueError("LZO1X: output overrun")

    def literals(k: int):
        nonlocal ip
        if ip + k > n:
            raise ValueError("LZO1X: input overrun")
        out.extend(src[ip:ip + k]); ip += k
        if max_out is not None and len(out) > max_out:
            raise ValueError("LZO1X: output overrun")

    def ext(base: int) -> int:
        nonlocal ip
        t = 0
        while True:
            if ip >= n:
                raise ValueError("LZO1X: input overrun")
            b = src[ip]; ip += 1
            if b:
                return t + base + b
            t += 255

    try:
        state = "loop"                     # "loop": a literal run may follow; "after_lit": right after a literal run; "match": t holds a match
        t = 0
        if n and src[0] > 17:
            t = src[0] - 17; ip = 1
            literals(t)
            state = "match_next0" if t < 4 else "after_lit"
        while True:
            if state in ("loop", "after_lit", "match_next0"):
                t = src[ip]; ip += 1
                if state == "loop" and t < 16:
                    k = (t + 3) if t else ext(15) + 3
                    literals(k)
                    state = "after_lit"
                    continue
                if state == "after_lit" and t < 16:          # 3-byte match just beyond the M2 window
                    copy_match(1 + 0x800 + (t >> 2) + (src[ip] << 2), 3); ip += 1
                    trail = src[ip - 2] & 3
                    if trail:
                        literals(trail); state = "match_next0"
                    else:
                        state = "loop"
                    continue
            # t is a match instruction (t >= 16, or an M1 after trailing literals)
            if t >= 64:
                copy_match(1 + ((t >> 2) & 7) + (src[ip] << 3), (t >> 5) - 1 + 2); ip += 1
            elif t >= 32:
                length = (t & 31) or ext(31)
                v = src[ip] | (src[ip + 1] << 8); ip += 2
                copy_match(1 + (v >> 2), length + 2)
            elif t >= 16:
                dist = (t & 8) << 11
                length = (t & 7) or ext(7)
                v = src[ip] | (src[ip + 1] << 8); ip += 2
                dist += v >> 2
                if dist == 0:
                    if length != 1:
                        raise ValueError("LZO1X: bad end marker")
                    break
                copy_match(dist + 0x4000, length + 2)
            else:                                               # M1: 2-byte match, only valid after trailing literals
                copy_match(1 + (t >> 2) + (src[ip] << 2), 2); ip += 1
            trail = src[ip - 2] & 3
            if trail:
                literals(trail); state = "match_next0"
            else:
                state = "loop"
    except IndexError:
        raise ValueError("LZO1X: input overrun") from None
    if ip != n:
        raise ValueError("LZO1X: trailing bytes after the end marker")
    return bytes(out)
