#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3 | cut -c1-300
