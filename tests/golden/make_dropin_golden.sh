#!/bin/sh
# Regenerates tests/golden/dropin/*.txt: the output of tests/cpp/dropin_check.cc compiled against the UNMODIFIED
# reference headers (oracle/_ref/dropin_check_ref, built by `make -C oracle ref` where /root/reference is mounted).
set -e
cd "$(dirname "$0")/../.."
make -C oracle ref >/dev/null
for c in readme_ring empty no_edges self_loop ring6 star ring100 random_full multi_equals_single string_keys long_keys; do
  ./oracle/_ref/dropin_check_ref $c > tests/golden/dropin/$c.txt
done
