timeout -k 10 300 python -m pytest tests/test_sharded_exchange.py -x -q -m gpu 2>&1 | tail -4
bash tools/_mg.sh 2 1
