#!/bin/bash
# A/B of the piece pipeline at a many-piece size: "name|library|ENV=..." lines of tools/ab_specs.txt
cd "$(dirname "$0")/.."
cp osmo-tetra_b200/libtetra_b200.so /tmp/keep.so
while IFS='|' read -r name lib envs; do
  [ -z "$name" ] && continue
  case "$name" in \#*) continue;; esac
  if [ "$lib" != "-" ]; then cp "$lib" osmo-tetra_b200/libtetra_b200.so; else cp /tmp/keep.so osmo-tetra_b200/libtetra_b200.so; fi
  printf "%-28s " "$name"; env $envs timeout 300 python tools/pipe_probe.py ${PROBE_N:-8e6} 2>&1 | tail -1
done < tools/ab_specs.txt
cp /tmp/keep.so osmo-tetra_b200/libtetra_b200.so
