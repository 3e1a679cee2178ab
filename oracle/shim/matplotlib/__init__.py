"""empty stub: the reference imports matplotlib but never plots"""
