#pragma once
#define BOOST_FOREACH(decl, col) for (decl : col)
