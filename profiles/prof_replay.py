import sys, json, argparse, torch
sys.path.insert(0, "/root/repo")
import bench
args = argparse.Namespace(no_cpu_baseline=False)
print(json.dumps(bench.bench_replay(args, torch.device("cuda:0"), 0, 1), indent=1))
