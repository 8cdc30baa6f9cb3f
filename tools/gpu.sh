#!/bin/bash
# rebuild the library (a stale .so is refused on the box), then run a command on a B200: tools/gpu.sh TIMEOUT 'command' [gpurun flags]
cd "$(dirname "$0")/.." && python -c "import __graft_entry__ as g; g.build()" > /tmp/build.log 2>&1 || { tail -20 /tmp/build.log; exit 1; }
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun $3 --timeout "$1" -- "$2"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 60
done
exit 3
