/* TEST INFRASTRUCTURE ONLY (oracle/): the reference includes an autoconf-generated config.h
   (rnacore/region.cc:9, rnacore/splice_graph.cc:9, meta/bundle.cc:8) but uses no macro from it. */
