timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 150 python -m pytest tests/test_gpu_reduce.py tests/test_gpu_radix.py tests/test_gpu_host_nodes.py -q 2>&1 | tail -3
