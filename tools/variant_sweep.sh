#!/bin/bash
# time warp_multi for every prebuilt variant in variants/ (tuning only)
cp deepvideocodec_b200/libdvc_b200.so /tmp/orig.so
for f in variants/lib_*.so; do
  cp $f deepvideocodec_b200/libdvc_b200.so
  for pf in 1 2; do
    echo -n "$(basename $f) "; DVC_WARP_PREFETCH=$pf python tools/warp_tune.py smooth | tail -1
  done
done
cp /tmp/orig.so deepvideocodec_b200/libdvc_b200.so
