"""Golden vectors of the reference for the hot path (tests/golden/kat.json, written
by tools/extract_golden.py from the reference tree) plus the edge inputs of
SURVEY.md appendix B."""
import json
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(_HERE, "golden", "kat.json")) as f:
    _K = json.load(f)

Q = 0x12AB655E9A2CA55660B44D1E5C37B00159AA76FED00000010A11800000000001

# tests/encoding.rs:61-78
GENERATOR_MULTIPLES = _K["generator_multiples"]
# src/ark_curve/elligator.rs:88-188
ELLIGATOR_INPUTS = _K["elligator_inputs"]
ELLIGATOR_XY = [(int(x), int(y)) for x, y in _K["elligator_xy"]]
# tests/encoding.proptest-regressions:7-9
REGRESSION_ENCODINGS = _K["regression_encodings"]

# (bytes, decodes?) -- SURVEY.md appendix B, each traced in the reference
EDGE_CASES = [
    (bytes(32), True),                                   # identity, tests/encoding.rs:19-26
    (bytes([8]) + bytes(31), True),                      # generator, tests/encoding.rs:28-52
] + [(bytes([b]) + bytes(31), False) for b in range(1, 8)] + [
    (bytes(REGRESSION_ENCODINGS[0]), True),              # [0,..,0,5]
    (bytes(REGRESSION_ENCODINGS[1]), True),              # [12,0,..]
    (bytes(REGRESSION_ENCODINGS[2]), False),             # >= q
    (bytes(31) + bytes([0x20]), False),                  # top bits set
    (bytes(31) + bytes([0x80]), False),
    ((1).to_bytes(32, "little"), False),                 # s = 1: u1 = 0
    ((Q - 1).to_bytes(32, "little"), False),             # s = -1: negative (q-1 is even? no: odd)
    (Q.to_bytes(32, "little"), False),                   # s = q: non-canonical zero
    (b"\xff" * 32, False),
]
EDGE_ENCODINGS = [b for b, _ in EDGE_CASES]
