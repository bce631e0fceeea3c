python -m pytest tests/test_gpu_u8_input.py -x -q 2>&1 | tail -5
python bench.py > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err; tail -c 2500 gpurun_out/s3_bench.json; tail -3 gpurun_out/s3_bench.err
