"""Per-call latency of the synchronous host API at the reference's batch-1 shapes (BASELINE configs 1-4)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import opencl_fft_b200 as eng  # noqa: E402

print(json.dumps({k: round(v, 2) for k, v in bench.bench_latency(eng, 0).items()}))
