#!/bin/sh
# Regenerates tests/golden/demo_reference.txt: the output of examples/demo_main.cc compiled against the UNMODIFIED reference
# headers on the reference's own dataset (/root/reference/example.txt, demo parameters of /root/reference/src/main.cc:37,50,64),
# and tests/golden/example_edges.csv.gz, the dataset as a fixture for the GPU box (where /root/reference does not exist).
# Takes a few minutes on the CPU container (the reference runs grankMulti, grank, mccompletepathv2 and 3 x 200 exact PPRs).
set -e
cd "$(dirname "$0")/../.."
REF=${REF:-/root/reference}
mkdir -p oracle/_ref
g++ -std=c++11 -O3 -march=x86-64-v3 -w -I$REF/include -I$REF/include/internal -I$REF/header-only -o oracle/_ref/demo_ref examples/demo_main.cc -lpthread
./oracle/_ref/demo_ref $REF/example.txt > tests/golden/demo_reference.txt
gzip -9 -c $REF/example.txt > tests/golden/example_edges.csv.gz
