#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 300 python scripts/latency_order.py 2>&1 | tail -4
