"""Development aid: build libot_b200_<tag>.so with extra -D flags for ONE source (A/B timing of kernel variants in one gpurun call).

    python tools/build_variant.py k1 ot_cdecoder.cu -DOT_CD_KHOIST=1
    OT_B200_LIB=$PWD/onnx-transformer_b200/build/libot_b200_k1.so python tools/decoder_trace.py --no-graph
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "onnx-transformer_b200"))
import build as B  # noqa: E402

tag, src, flags = sys.argv[1], sys.argv[2], sys.argv[3:]
B.build()
obj = os.path.join(B.BUILD, "var_%s.o" % tag)
subprocess.check_call([B._nvcc(), *B.NVCC_FLAGS, *flags, "-c", os.path.join(B.CSRC, src), "-o", obj])
objs = [os.path.join(B.BUILD, f[:-3] + ".o") for f in sorted(os.listdir(B.CSRC)) if f.endswith(".cu") and f != src] + [obj]
out = os.path.join(B.BUILD, "libot_b200_%s.so" % tag)
subprocess.check_call([B._nvcc(), "-shared", "-o", out, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"])
print(out)
