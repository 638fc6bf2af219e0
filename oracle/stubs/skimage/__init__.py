# Minimal stand-in for scikit-image (absent here); only measure.block_reduce has a real body.
