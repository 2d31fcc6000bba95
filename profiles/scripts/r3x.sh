# the round's last call: full GPU suite on the final tree
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r3x_pytest.log; tail -3 gpurun_out/r3x_pytest.log
