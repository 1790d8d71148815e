#!/bin/bash
cd ${GRAFT_REPO_ROOT:-.}
bash scripts/sanitize.sh host 2>&1 | tail -45
bash scripts/r2_profile.sh
