#!/bin/bash
timeout 300 python tools/pcie2d_probe.py 2>&1 | tail -14
