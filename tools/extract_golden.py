#!/usr/bin/env python3
"""Extract the reference's golden vectors for the hot path into tests/golden/kat.json.

Run in the authoring container (needs /root/reference); the JSON it writes is
committed so that nothing at test time reads the reference tree.

Sources (reference crate v0.10.1):
  tests/encoding.rs:61-78              16 encodings of i*G
  src/ark_curve/elligator.rs:88-188    8 Elligator inputs and expected affine (x, y)
  tests/encoding.proptest-regressions:7-9  three saved edge encodings
  proptest-regressions/invsqrt.txt:7   sqrt_ratio_zeta(1, 1)
"""
import json
import re
import sys
from pathlib import Path

REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "kat.json"

enc_rs = (REF / "tests/encoding.rs").read_text()
block = enc_rs[enc_rs.index("let expected_points = ["):]
block = block[:block.index("];")]
gen_multiples = re.findall(r'"([0-9a-f]{64})"', block)
assert len(gen_multiples) == 16

ell = (REF / "src/ark_curve/elligator.rs").read_text()
ib = ell[ell.index("let inputs = ["):ell.index("let expected_xy_coordinates")]
inputs = [[int(x) for x in re.findall(r"\d+", grp)] for grp in re.findall(r"\[\s*((?:\d+,\s*)+\d+,?\s*)\]", ib)]
assert len(inputs) == 8 and all(len(v) == 32 for v in inputs)
xb = ell[ell.index("let expected_xy_coordinates"):ell.index("use ark_serialize::CanonicalDeserialize")]
coords = re.findall(r'"(\d{60,80})"', xb)
assert len(coords) == 16
xy = [[coords[2 * i], coords[2 * i + 1]] for i in range(8)]

reg = (REF / "tests/encoding.proptest-regressions").read_text()
edge = [[int(x) for x in re.findall(r"\d+", m)] for m in re.findall(r"bytes = \[([^\]]+)\]", reg)]
assert len(edge) == 3 and all(len(v) == 32 for v in edge)

OUT.parent.mkdir(parents=True, exist_ok=True)
OUT.write_text(json.dumps({
    "source": "penumbra-zone/decaf377 v0.10.1",
    "generator_multiples": gen_multiples,
    "elligator_inputs": inputs,
    "elligator_xy": xy,
    "regression_encodings": edge,
}, indent=1) + "\n")
print("wrote", OUT)
