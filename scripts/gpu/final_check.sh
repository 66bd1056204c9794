#!/bin/bash
bash scripts/gpu/suite.sh
bash scripts/gpu/bench_and_profiles.sh
