"""(empty: the reference only imports this module)"""
