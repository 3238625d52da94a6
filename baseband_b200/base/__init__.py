"""Format-independent host layer (mirrors the reference's ``baseband.base``)."""
