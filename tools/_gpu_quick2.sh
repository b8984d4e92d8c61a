timeout 600 python tools/bench_config5.py > gpurun_out/config5.json 2>gpurun_out/config5.err; cat gpurun_out/config5.json; tail -3 gpurun_out/config5.err
