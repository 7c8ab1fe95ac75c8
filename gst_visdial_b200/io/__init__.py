"""Host-side input / output formats around the hot path."""
