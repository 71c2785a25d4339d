python -m pytest tests/test_train_gpu.py -q -m gpu -x 2>&1 | tail -3
POEM_TRAIN_PROF=1 timeout -s KILL 300 python scripts/bench_train.py medium 8 32 2>&1 | head -32
