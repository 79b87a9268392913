"""Shared helpers for the parity tests (oracle side lives in oracle/)."""
import numpy as np

from oracle import decaf377_ref as o


def np_bytes(chunks, width):
    return np.frombuffer(b"".join(chunks), np.uint8).reshape(-1, width).copy()


def mont(vals):
    return np_bytes([o.fq_to_mont_bytes(v) for v in vals], 32)


def unmont(arr):
    return [o.fq_from_mont_bytes(arr[i].tobytes()) for i in range(arr.shape[0])]


def canon(vals, mod_bytes=32):
    return np_bytes([int(v).to_bytes(32, "little") for v in vals], 32)


def wire(points):
    return np_bytes([o.point_to_wire(p) for p in points], 128)


def unwire(arr):
    return [o.point_from_wire(arr[i].tobytes()) for i in range(arr.shape[0])]


def oracle_points(tag, n):
    """P_i = encode_to_curve(from_le_bytes_mod_order(B(tag, i)))  (tests/operations.rs:6-11)."""
    return [o.encode_to_curve(o.fq_from_le_bytes_mod_order(b)) for b in o.xof_blocks(tag, n)]


def oracle_scalars(tag, n):
    return [o.fr_from_le_bytes_mod_order(b) for b in o.xof_blocks(tag, n)]
