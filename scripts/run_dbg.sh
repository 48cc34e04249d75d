for d in ${2:-0 1 2 3}; do echo "== PCFD_WS_DEBUG=$d"; PCFD_WS_DEBUG=$d timeout 120 python scripts/bench_layers.py --engines 2 --passes fwd --only "${1:-int L}" 2>&1 | grep -E "engine 2"; done
