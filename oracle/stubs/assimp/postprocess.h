// stub, see Importer.hpp
