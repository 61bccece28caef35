#!/bin/bash
for r in 4 8 16 32; do echo "== SEG_ROWS=$r"; ELAS_B200_SEG_ROWS=$r tools/ncu_launches.sh seg$r | grep -E "k_seg|sum"; done
