#!/bin/bash
# Generates tests/golden/ref_*.json from the UNMODIFIED reference.  Needs what this image does not have: cargo and
# network access (crates.io + github.com/sphereflow/collision2d at the revision Cargo.lock pins).
#   rust/dump_golden/run.sh /path/to/light_garden_checkout
# 1. writes the scene files (RON) of the small parity specs:   tests/golden/ref_scenes/*.ron
# 2. copies dump_golden.rs into the checkout as src/bin/dump_golden.rs (the only file added; nothing is modified)
# 3. cargo run --release --locked --bin dump_golden -- tests/golden  <scenes>
# After that `python -m pytest tests/test_reference_golden.py` stops skipping and holds oracle/ to the vectors.
set -euo pipefail
REPO="$(cd "$(dirname "$0")/../.." && pwd)"
REF="${1:?path to a checkout of sphereflow/light_garden}"
command -v cargo >/dev/null || { echo "cargo not found: this step needs a Rust toolchain" >&2; exit 3; }
python "$REPO/tests/golden/make_ref_scenes.py"
mkdir -p "$REF/src/bin"
cp "$REPO/rust/dump_golden/dump_golden.rs" "$REF/src/bin/dump_golden.rs"
cp "$REF/default.ron" "$REPO/tests/golden/ref_scenes/default.ron"
(cd "$REF" && cargo run --release --locked --bin dump_golden -- "$REPO/tests/golden" "$REPO"/tests/golden/ref_scenes/*.ron)
rm -f "$REF/src/bin/dump_golden.rs"
ls -la "$REPO"/tests/golden/ref_*.json
