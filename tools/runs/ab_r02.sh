echo "== pt_probe v4 (TMA key stream)"; python tools/ncu_join.py 64 0x80 | tail -1; python tools/ncu_join.py 8 0x80 | tail -1; python tools/ncu_join.py 32 0x80 1.25 | tail -1
