for M in 38400 75776 1048576; do echo "--- M=$M"; timeout 60 python tools/bench_mlp.py $M train 2>&1 | tail -1 | cut -c1-150; done
timeout 400 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
