# ORACLE -- test infrastructure only.  Compiles the reference's OWN sources, where they lie under $(REF) (nothing is copied
# into this repository), for the two translation units of the hot path's neighbourhood that need no third-party library:
#   src/datastructure/Interval.cpp   (hrm::Interval::unions / intersects / complements: the interval algebra of K3)
#   src/util/Parse2dCsvFile.cpp      (hrm::parse2DCsvFile: every scene / robot / end-point file goes through it)
# Everything else on the path needs Eigen >= 3.3 (absent from this image, no network): see DESIGN.md.
# Flags = the reference's Release build (CMakeLists.txt:6-11: -O3 -DNDEBUG, plain x86-64, no FMA contraction).
# Output only into oracle/_ref/ (git-ignored, shipped to the GPU box with the repo snapshot).
#     make -f oracle/ref_build.mk            (from the repository root; REF=/root/reference)
REF ?= /root/reference
CXX := g++
OUT := oracle/_ref
all: $(OUT)/libhrm_ref.so
$(OUT)/libhrm_ref.so: $(REF)/src/datastructure/Interval.cpp $(REF)/src/util/Parse2dCsvFile.cpp oracle/ref_shim.cpp oracle/ref_stubs/eigen3/Eigen/Dense
	mkdir -p $(OUT)
	$(CXX) -O3 -DNDEBUG -std=c++17 -fPIC -ffp-contract=off -shared -I$(REF)/include -Ioracle/ref_stubs -o $@ \
	    $(REF)/src/datastructure/Interval.cpp $(REF)/src/util/Parse2dCsvFile.cpp oracle/ref_shim.cpp
.PHONY: all
