# 2-GPU validation of the whole GPU suite
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -12
