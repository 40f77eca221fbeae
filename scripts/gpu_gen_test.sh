python -c "import __graft_entry__ as e; e.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_sampler.py tests/test_generator_parity.py -m gpu -x -q 2>&1 | tail -40
