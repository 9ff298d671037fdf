# usage: tools/flake_probe.sh [runs]  -- repeats the whole GPU suite and reports every failure (tests that depend on float-atomic order)
n=${1:-8}; f=0
for i in $(seq 1 $n); do
  python -m pytest tests -m gpu -q --tb=line 2>&1 | grep -E "passed|failed|FAILED" | cut -c1-200 | tee /tmp/flake_$i.txt | tail -3
  grep -q failed /tmp/flake_$i.txt && f=$((f+1))
done
echo "$f of $n suite runs had a failure"
