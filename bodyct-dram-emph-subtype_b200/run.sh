#!/bin/sh
# Same entry point as the reference's run.sh (processor.py with its defaults: /input -> /output).
exec python3 "$(dirname "$0")/processor.py" "$@"
