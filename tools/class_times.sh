#!/bin/bash
# warm-cache device time of every launch class of one sweep (development aid)
for c in 0 1 2 3 4 5; do
  echo -n "class $c: "; GSG_ONLY_CLASS=$c python tools/probe.py --steps 1 --reps 5 2>&1 | grep -E "apply D_1 beta=0|apply D_6 beta=1" | tr '\n' ' '; echo
done
