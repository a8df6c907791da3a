#!/bin/bash
for nb in 999 500 250 100; do echo "== nbins $nb"; python tools/profile_pair.py c2 214 3 $nb | tail -1; done
for cap in 800 1200; do echo "== nbins 999 cap $cap"; AMOFB_TILE_CAP=$cap python tools/profile_pair.py c2 214 3 999 | tail -1; done
