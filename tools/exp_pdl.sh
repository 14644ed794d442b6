#!/bin/bash
for i in 1 2; do
echo "== graph, PDL off"; HRP_SWEEP=512:1 python tools/profile_model.py sweep
echo "== graph, PDL on"; HRP_PDL=1 HRP_SWEEP=512:1 python tools/profile_model.py sweep
echo "== eager, PDL off"; HRP_NO_GRAPH=1 HRP_SWEEP=512:1 python tools/profile_model.py sweep
echo "== eager, PDL on"; HRP_NO_GRAPH=1 HRP_PDL=1 HRP_SWEEP=512:1 python tools/profile_model.py sweep
done
