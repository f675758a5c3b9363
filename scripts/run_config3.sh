#!/bin/bash
# BASELINE config 3 at full size: 4-way chain join with range/equality filters and a 3-column SUM projection
# over a 200M-row fact relation, through the reference's own ExecuteQuery (query.o + JoinEnum):
#   oracle/_ref/ref_driver  = the unmodified reference operators (CPU, 8 threads: at 16 the reference loses tuples)
#   oracle/_ref/b200_driver = the same host objects over libb200join.so (link-time drop-in)
# prints both result lines (must be identical) and the per-repetition wall-clock seconds.
set -e
cd "$(dirname "$0")/.."
F=${1:-200000000}; D1=$((1<<24)); D2=$((1<<20)); D3=$((1<<16))
Q="0 1 2 3|0.1=1.0&1.1=2.0&2.1=3.0&0.3>2499&0.3<7500&0.4=1|0.2 1.2 3.1"
SPECS="synth:$F:iota,uni$D1@11,pay@12,uni10000@13,uni4@14 synth:$D1:iota,uni$D2@21,pay@22 synth:$D2:iota,uni$D3@31,pay@32 synth:$D3:iota,pay@41"
echo "== reference (CPU)"; oracle/_ref/ref_driver -t 8 -r 2 $SPECS -- "$Q"
echo "== drop-in (B200)"; oracle/_ref/b200_driver -t 1 -r 4 $SPECS -- "$Q"
