#!/bin/bash
# time kernel variants: each .so in profiles/variants plus the default build
python profiles/quick_time.py 2>&1 | grep "n=67108864" | sed "s/^/default: /"
for so in profiles/variants/*.so; do
  B2048_LIB=$PWD/$so python profiles/quick_time.py 2>&1 | grep "n=67108864" | sed "s#^#$so: #"
done
