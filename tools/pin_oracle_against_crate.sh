#!/bin/bash
# Turns "parity unpinned at the bit level" into pinned, on any machine that has cargo (this repository's build image has not):
# runs the REAL differential-equations crate on a fixed list of cases and stores every output as f64 bit patterns;
# tests/test_oracle_golden.py::test_oracle_matches_the_crate_bit_for_bit then replays them through oracle/oracle.cpp.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
command -v cargo >/dev/null || { echo "cargo not found: nothing pinned (this is the state of the build image)"; exit 3; }
cd "$ROOT/oracle/crate_pin"
cargo run --release > "$ROOT/tests/golden/reference_bits.json"
echo "wrote tests/golden/reference_bits.json; now run: python -m pytest tests/test_oracle_golden.py -k crate_bit_for_bit"
