#!/bin/bash
set -u
for T in 1 160; do
for D in 0 1 2 3 4 8 9 15; do
  IQ_CHAIN_DBG=$D timeout 60 python scripts/chain_probe.py 64 64 128 32 $T 2>&1 | tail -1
done
done
IQ_CHAIN_DBG=0 timeout 60 python scripts/chain_probe.py 64 32 64 32 1 2>&1 | tail -1
IQ_CHAIN_DBG=0 timeout 60 python scripts/chain_probe.py 32 64 128 32 1 2>&1 | tail -1
