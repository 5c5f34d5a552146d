python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python tools/bench_configs.py 32768 2>&1 | tail -9
