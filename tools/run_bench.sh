cp webgraph-ans-rs_b200/libwgans.so /tmp/base.so
python tools/prof_decode.py eu-2015-host-shaped 3 2>&1 | tail -1
for t in k2_blocks=1184 k2_blocks=1332; do echo $t; WGA_TUNING=$t python tools/prof_decode.py eu-2015-host-shaped 3 2>&1 | tail -1; done
for v in spv1 spv3; do cp build_variants/libwgans_$v.so webgraph-ans-rs_b200/libwgans.so; echo $v; python tools/prof_decode.py eu-2015-host-shaped 3 2>&1 | tail -1; done
cp /tmp/base.so webgraph-ans-rs_b200/libwgans.so
