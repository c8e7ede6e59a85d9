#!/bin/bash
# compute-sanitizer over the small end-to-end invocation (smoke()): memcheck, then racecheck + synccheck
# of the shared-memory kernels. Run on the GPU box; takes a few minutes (the tools serialise kernels).
set -u
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
    timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 \
        python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_$tool.log 2>&1
    echo "$tool rc=$?"
    tail -n 3 gpurun_out/sanitize_$tool.log
done
