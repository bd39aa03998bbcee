python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python tools/exp_single.py cfg3 2>&1 | grep -E "single-walk|checksums|hybrid" | head -8
