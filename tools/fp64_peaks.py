"""Measure the fp64 denominators of this GPU (SURVEY.md section 8d asks for them next to MEASURED_PEAKS.json's HBM figure):
sustained fp64 FMA rate of the CUDA cores, sustained mma.sync.m8n8k4.f64 rate, cuBLAS DGEMM.  Writes one JSON object.
usage: python tools/fp64_peaks.py [out.json]"""
import json
import sys
sys.path.insert(0, ".")
import hybridsbp_b200 as hs

ctx = hs.Context(0)
pk = ctx.fp64_peaks(8192)
out = {"fp64_fma_tflops": pk["fma"], "fp64_dmma_mma_sync_tflops": pk["dmma"], "dgemm_8192_tflops": pk["dgemm"],
       "how": "k_peak_fma: 8 FMA chains per thread, 256 threads x 8 CTAs per SM, 20000 rounds; k_peak_dmma: 4 independent "
              "mma.sync.m8n8k4.f64 accumulators per warp, same grid; cuBLAS cublasDgemm 8192^3; best of 5, CUDA events"}
print(json.dumps(out))
if len(sys.argv) > 1:
    with open(sys.argv[1], "w") as f:
        json.dump(out, f, indent=1)
