"""`dynamicemb.benchmark` — import path of the reference's synthetic id generators (corelib/dynamicemb/benchmark/)."""
