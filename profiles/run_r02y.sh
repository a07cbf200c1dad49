mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_integration_adapters.py tests/test_episode_writer.py tests/test_trajectories.py -x -q -m gpu > gpurun_out/r02y_tests.log 2>&1; tail -8 gpurun_out/r02y_tests.log
