#!/bin/bash
for kb in 0 16 32; do echo "== extra smem ${kb} KB"; UPSNET_DCN_EXTRA_SMEM_KB=$kb UPSNET_DCN_TILE=0 python scripts/dcn_tile_exp.py run; done
