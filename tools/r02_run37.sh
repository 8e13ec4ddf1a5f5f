#!/bin/bash
# where the 0.58 ms of the direct-download render go: warp shape (16 x 2 vs 8 x 4) x destination (resident / pinned host images)
echo "== resident, default (8x4)"; python tools/tiny_probe_ms.py bunny4k
echo "== resident, 16x2"; CUTRACE_DEBUG_WIDE_WARPS=1 python tools/tiny_probe_ms.py bunny4k
echo "== direct download, default (16x2)"; python tools/e2e_probe.py bunny4k 2>&1 | tail -2
echo "== direct download, 8x4"; CUTRACE_DEBUG_WIDE_WARPS=0 python tools/e2e_probe.py bunny4k 2>&1 | tail -2
