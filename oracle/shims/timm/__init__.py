"""Import shim so the read-only reference's `models/__init__.py` can be imported in the
build container (it pulls in Swin/window-attention files that need three timm symbols).
Test infrastructure only."""
