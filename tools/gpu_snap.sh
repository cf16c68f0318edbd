#!/bin/bash
# Run a command on the GPU box against a FROZEN copy of the tree (.snap/<name>), so that the working tree can keep
# changing while the call waits for a GPU slot (gpurun snapshots /root/repo only when it gets the box).
#   tools/gpu_snap.sh <name> <timeout_s> [--gpus N] -- '<command run from the root of the frozen copy>'
# Whatever the command writes under gpurun_out/ comes back into /root/repo/gpurun_out/.
set -u
name=$1; shift
tmo=$1; shift
gp=""
if [ "$1" = "--gpus" ]; then gp="--gpus $2"; shift 2; fi
[ "$1" = "--" ] && shift
cmd=$1
root=/root/repo
snap=$root/.snap/$name
rm -rf "$snap"; mkdir -p "$snap"
tar -C "$root" --exclude=./.git --exclude=./gpurun_out --exclude=./.snap --exclude=__pycache__ --exclude=.pytest_cache --exclude='*.o' -cf - . | tar -C "$snap" -xf -
full="cd .snap/$name && rm -rf gpurun_out && ln -s \$GRAFT_REPO_ROOT/gpurun_out gpurun_out && $cmd"
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$tmo" $gp -- "$full"
  rc=$?
  if [ $rc -ne 3 ]; then rm -rf "$snap"; exit $rc; fi
  sleep 45
done
rm -rf "$snap"
exit 3
