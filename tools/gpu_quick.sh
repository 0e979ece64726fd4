#!/bin/bash
timeout 600 python tools/two_stream_probe.py 2>&1 | tail -12
