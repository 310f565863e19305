#!/bin/bash
set -u
for rep in 1 2 3; do
timeout 200 python tools/graph_modes.py --captures 4 2>/dev/null | grep capture | sed "s/^/proc $rep /"
done
