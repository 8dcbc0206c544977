timeout 200 python tools/probe_halo_speed.py 2>&1 | tail -12
DVD_CONV_HALO=1 timeout 200 python -m pytest tests/test_conv2d_gpu.py -q -k "not stream_k" 2>&1 | tail -4
