set -x
timeout 200 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
