// Empty stand-in: the reference header that includes this only needs the name to
// resolve; none of the functions compiled for the oracle use anything from Boost here.
// (Boost is not installed in this image; see oracle/README.md.)
#pragma once
