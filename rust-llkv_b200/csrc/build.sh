#!/bin/bash
# Builds libllkv_gpu.so (sm_100a) next to the sources.  nvcc cross-compiles without a GPU.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall"
mkdir -p build
# the lean kernel's sources, embedded for run-time specialisation (jit.cpp)
python3 - <<'PY'
import os
out = []
for name in ("lean_kernel.cuh", "device_util.cuh", "plan.h"):
    text = open(name).read()
    assert ')LLKVSRC"' not in text
    ident = "kSrc_" + name.replace(".", "_")
    # string literals are capped at 64 KiB by some compilers: emit chunks
    chunks = [text[i:i + 12000] for i in range(0, len(text), 12000)]
    out.append("static const char %s[] =\n%s;\n" % (ident, "\n".join('R"LLKVSRC(%s)LLKVSRC"' % c for c in chunks)))
new = "".join(out)
path = "build/lean_sources.inc"
if not os.path.exists(path) or open(path).read() != new:
    open(path, "w").write(new)
PY
pids=()
$NVCC $FLAGS -c scan_kernel.cu -o build/scan_kernel.o & pids+=($!)
$NVCC $FLAGS -c llkv_gpu.cu -o build/llkv_gpu.o & pids+=($!)
$NVCC $FLAGS ${LLKV_PTXAS_V:+-Xptxas -v} -c lean_kernel.cu -o build/lean_kernel.o & pids+=($!)
$NVCC $FLAGS ${LLKV_PTXAS_V:+-Xptxas -v} -c partition_kernel.cu -o build/partition_kernel.o & pids+=($!)
g++ -O2 -std=c++17 -fPIC -Wall -Wno-nonnull -c compiler.cpp -o build/compiler.o & pids+=($!)
g++ -O2 -std=c++17 -fPIC -Wall -I/usr/local/cuda/include -c jit.cpp -o build/jit.o & pids+=($!)
g++ -O2 -std=c++17 -fPIC -Wall -c descriptor.cpp -o build/descriptor.o & pids+=($!)
g++ -O3 -std=c++17 -fPIC -Wall -I/usr/local/cuda/include -c upload.cpp -o build/upload.o & pids+=($!)
for pid in "${pids[@]}"; do wait "$pid"; done  # any failed compile fails the build (set -e)
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o libllkv_gpu.so build/scan_kernel.o build/lean_kernel.o build/partition_kernel.o build/llkv_gpu.o build/compiler.o build/jit.o build/descriptor.o build/upload.o -cudart static -ldl -lpthread -lrt
echo built $(pwd)/libllkv_gpu.so
