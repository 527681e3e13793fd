#!/bin/bash
python tools/gp_probe.py 20 2>&1 | grep -v Warn | grep -v "kernel"
