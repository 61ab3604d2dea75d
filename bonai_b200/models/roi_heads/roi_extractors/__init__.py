from .single_level import SingleRoIExtractor, BaseRoIExtractor

__all__ = ['SingleRoIExtractor', 'BaseRoIExtractor']
