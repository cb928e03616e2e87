"""Host-side corridor geometry, mirroring the reference's callers of SingleAlign.

These builders stay host code in the reference; they are restated
here in float32 numpy so that tests and bench.py can generate exactly the `CorridorLine[]`
arrays ngmlr would hand to `IAlignment::SingleAlign`:

  corridor_endpoints_with_anchors  <- AlignmentBuffer::getCorridorEndpointsWithAnchors
                                      (src/AlignmentBuffer.cpp:129-197)
  corridor_endpoints               <- getCorridorEndpoints   (src/AlignmentBuffer.cpp:107-127)
  corridor_linear                  <- getCorridorLinear      (src/AlignmentBuffer.cpp:68-82)
  corridor_full                    <- getCorridorFull        (src/AlignmentBuffer.cpp:84-105)

All return (offsets int32[H], lengths int32[H]). Float math is float32 with C truncation.
"""
import numpy as np

F = np.float32


def _trunc(a):
    return np.asarray(a).astype(np.int32)  # C float->int conversion truncates toward zero


def corridor_endpoints_with_anchors(qry_len, ref_len, anchors_x, anchors_y, multiplier=1):
    """anchors_x / anchors_y: anchor positions relative to the interval start (ref) and to the
    aligned read part (read), as computed at src/AlignmentBuffer.cpp:149-158."""
    k = F(qry_len) * F(1.0) / F(ref_len)
    left = F(0.0)
    right = F(0.0)
    ax = np.asarray(anchors_x, dtype=np.int64)
    ay = np.asarray(anchors_y, dtype=np.int64)
    if ax.size:
        x_found = ax.astype(np.float32)
        x_expect = (ay.astype(np.float32) - F(0.0)) / k
        diff = (x_expect - x_found).astype(np.float32)
        pos = diff[diff > 0]
        neg = diff[~(diff > 0)]
        if pos.size:
            right = max(right, F(pos.max()))
        if neg.size:
            left = max(left, F((neg * F(-1.0)).max()))
    left = F(left + F(128))
    right = F(right + F(128))
    left = F(left + F(F(left + right) * F(0.1)))
    right = F(right + F(F(left + right) * F(0.1)))
    left = F(left * F(multiplier))
    right = F(right * F(multiplier))
    width = int(_trunc(F(left + right)))
    i = np.arange(qry_len, dtype=np.float32)
    offs = _trunc(((i - F(0.0)) / k).astype(np.float32) - right)
    lens = np.full(qry_len, width, dtype=np.int32)
    return offs, lens


def corridor_endpoints(qry_len, ref_len, corridor, realign=False):
    width = corridor // (1 if realign else 4)
    k = F(qry_len) * F(1.0) / F(ref_len)
    d = F(width) / F(2.0)
    i = np.arange(qry_len, dtype=np.float32)
    offs = _trunc(((i - d) / k).astype(np.float32))
    return offs, np.full(qry_len, width, dtype=np.int32)


def corridor_linear(qry_len, corridor):
    i = np.arange(qry_len, dtype=np.int32)
    return (i - corridor // 2).astype(np.int32), np.full(qry_len, corridor, dtype=np.int32)


def corridor_full(qry_len, ref_len):
    off = int(np.float64(ref_len) * -0.2)  # (int)(corridorWidth * -0.2): double math
    length = ref_len + int(np.float64(ref_len) * 0.2)
    return (np.full(qry_len, off, dtype=np.int32), np.full(qry_len, length, dtype=np.int32))


def estimate_corridor(on_read, on_ref):
    """AlignmentBuffer::estimateCorridor, src/AlignmentBuffer.cpp:1454-1467.
    on_read / on_ref: interval lengths on read and reference."""
    diff = on_read - on_ref
    from_diff = int(F(abs(diff)) * F(2.1))
    from_len = int(F(abs(on_read)) * F(0.20))
    return min(8192, max(from_diff, from_len))
