#!/bin/bash
for t in 4 2 1; do echo "TABLES_PER_PASS=$t"; DN_GP_TABLES_PER_PASS=$t python tools/gp_probe.py 20 2>&1 | grep "tables"; done
