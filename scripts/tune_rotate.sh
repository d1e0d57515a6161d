for r in 5 0 3 7; do echo "== rotate $r"; FQ_TILE_ROTATE=$r timeout 200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>/dev/null | cut -c1-200; done
timeout 300 python -m pytest tests -m gpu -x -q -k "tile_fused or full_size" 2>&1 | tail -2
