#!/bin/bash
# sanitize.sh [outdir] : compute-sanitizer memcheck / racecheck / initcheck / synccheck over profiles/sanitize_run.py
out=${1:-gpurun_out}
for tool in memcheck racecheck initcheck synccheck; do
  extra=""
  [ $tool = memcheck ] && extra="--leak-check full"
  [ $tool = racecheck ] && extra="--racecheck-report all"
  timeout 1500 compute-sanitizer --tool $tool $extra --print-limit 50 python profiles/sanitize_run.py > $out/sanitizer_$tool.txt 2>&1
  echo "$tool rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|sanitize_run: done' $out/sanitizer_$tool.txt | tr '\n' ' ')"
done
