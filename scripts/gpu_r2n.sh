#!/bin/bash
set -x
O=gpurun_out/r2n
mkdir -p $O
python -c "import glass_b200.build as b; print(b.build())" > $O/build.log 2>&1
for i in 1 2; do
timeout 1800 python -m pytest tests -m gpu -q --timeout=900 > $O/pytest_$i.log 2>&1
echo "pytest rc $?" >> $O/pytest_$i.log
done
tail -n 4 $O/pytest_*.log
