python examples/train_tower.py --tower-height 2 --envs 1024 --iters 12 --log gpurun_out/s24_train_h2.jsonl > gpurun_out/s24_train_h2.out 2>&1; tail -2 gpurun_out/s24_train_h2.out
python examples/train_tower.py --tower-height 4 --max-steps 15 --envs 1024 --iters 30 --log gpurun_out/s24_train_h4.jsonl > gpurun_out/s24_train_h4.out 2>&1; tail -2 gpurun_out/s24_train_h4.out
